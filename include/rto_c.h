/* rto_c.h -- C ABI of the B200-native ray-casting core (librto.so).
 *
 * This is the drop-in boundary for the reference's ray-casting hot path.  The reference
 * (abodthedude25/Ray_Tracing_Octrees) has no FFI layer: the path is reached through C++ class calls
 *   BVH::BVH / BVH::query                              453-skeleton/BVH.h:44-63, BVH.cpp:19-113
 *   createOctreeFromVoxelGrid / freeOctree             453-skeleton/OctreeVoxel.h:66-69, OctreeVoxel.cpp:765-778
 *   RayTracerBVH::setOctree / renderSceneCompute       453-skeleton/RayTracerBVH.h:30-50, RayTracerBVH.cpp:430-505, 614-704
 *   static octreeRaySkip (VolumeRaycastRenderer)       453-skeleton/VolumeRaycastRenderer.cpp:50-155
 *   MarchingCubesRenderer::render / localMC            453-skeleton/Renderer.cpp:14-36, OctreeVoxel.cpp:780-879
 *   Camera::getView / getPos                           453-skeleton/Camera.cpp:11-29
 * Each entry point below names the reference interface it replaces.  C++ shims that keep the
 * reference's class names on top of this ABI live in ray_tracing_octrees_b200/csrc/shim/.
 *
 * Conventions: plain pointers and sizes only; every function returns an RtoStatus (0 = OK) unless
 * stated; nothing throws across the boundary; rto_last_error() gives a thread-local message.
 * All ray work runs on the GPU (sm_100a).  There is NO CPU fallback: without a usable CUDA device
 * every rto_scene_* / rto_render* / rto_trace* call fails with RTO_ERR_NO_DEVICE.
 * Host-side builders (rto_host_*) are CPU code exactly where the reference's are (tree construction).
 */
#ifndef RTO_C_H
#define RTO_C_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTO_API __attribute__((visibility("default")))

typedef enum RtoStatus {
	RTO_OK = 0,
	RTO_ERR_INVALID = 1,      /* bad argument */
	RTO_ERR_NO_DEVICE = 2,    /* CUDA device missing / wrong architecture */
	RTO_ERR_CUDA = 3,         /* CUDA runtime error; see rto_last_error() */
	RTO_ERR_ALLOC = 4,
	RTO_ERR_IO = 5,
	RTO_ERR_UNSUPPORTED = 6
} RtoStatus;

/* ---- reference-compatible plain data ------------------------------------------------------------- */

/* GPUNodes, RayTracerBVH.h:21-26: 15 x int32, std430 stride 60.  Index in the array == node id. */
typedef struct RtoGpuNode {
	int32_t x, y, z, size;
	int32_t isLeaf, isSolid, isUniform;
	int32_t child[8];
} RtoGpuNode;

/* Triangle, BVH.h:7-11: 9 floats (v0, v1, v2), 36 bytes. */
typedef struct RtoTriangle { float v0[3], v1[3], v2[3]; } RtoTriangle;

/* Per-frame camera constants.  The GLSL generateRay (RayTracerBVH.cpp:338-355) recomputes inverse(view)
 * and tan(fov/2) per pixel; here they are computed once on the host (rto_host_camera_orbit, or by the
 * caller with glm) so libm never runs on the GPU.  invView is column-major like glm::mat4. */
typedef struct RtoCamera {
	float   camPos[3];
	float   invView[16];
	float   tanHalfFov;
	float   aspect;
	int32_t width, height;
} RtoCamera;

/* Output planes of one render call.  Any pointer may be NULL (plane not wanted).  Row 0 is the top row
 * (gl_GlobalInvocationID.y == 0).  Pointers are host or device addresses according to `memory`. */
typedef enum RtoMemory { RTO_MEM_HOST = 0, RTO_MEM_DEVICE = 1 } RtoMemory;
typedef struct RtoFrame {
	float*   rgba;     /* 4 floats / pixel: Lambert colour of RayTracerBVH.cpp:331-336, alpha 1, miss = 0,0,0,1 */
	int32_t* hitId;    /* BVH scenes: index of the hit triangle in the caller's array; octree scenes: node index
	                      (== leaf id of RayTracerBVH::setOctree's BFS numbering); -1 on miss */
	float*   t;        /* hit distance along the (unit) ray; 1e30f on miss */
	int32_t  memory;   /* RtoMemory */
} RtoFrame;

/* Traversal semantics selector. */
typedef enum RtoMode {
	RTO_MODE_BVH = 0,           /* BVH::query candidate set + Moller-Trumbore closest hit (SURVEY.md 8c rule) */
	RTO_MODE_OCTREE_SKIP = 1,   /* octreeRaySkip semantics, VolumeRaycastRenderer.cpp:50-155 ("mode A") */
	RTO_MODE_OCTREE_GLSL = 2    /* intersectOctreeIterative semantics, RayTracerBVH.cpp:239-327 ("mode B") */
} RtoMode;

#define RTO_FLAG_SHADOWS      1u   /* BVH scenes: one shadow ray per primary hit towards normalize(1,1,1) */
#define RTO_FLAG_NO_PRUNE     2u   /* BVH scenes: visit every box the reference's queryNode visits (no t-pruning) */
#define RTO_FLAG_SORT_RAYS    4u   /* rto_trace_rays: trace the list in coherence order (octant, Morton cell of the entry point into the
                                      scene box, coarse direction; radix sort on the device); results are the same, in the caller's order */

typedef struct RtoScene RtoScene;          /* device-resident scene (BVH or octree), one CUDA stream each */
typedef struct RtoHostBvh RtoHostBvh;      /* host BVH identical in shape to the reference's BVH */

/* ---- library ----------------------------------------------------------------------------------------- */
RTO_API const char* rto_version(void);
RTO_API const char* rto_last_error(void);
/* Select the CUDA device for this thread (cudaSetDevice) and check it is sm_100. */
RTO_API int rto_init(int device);
RTO_API int rto_device_info(int* smCount, int* ccMajor, int* ccMinor, size_t* l2Bytes, size_t* totalMem);

/* ---- host-side builders (CPU, like the reference's own) ---------------------------------------------- */

/* createOctreeFromVoxelGrid (OctreeVoxel.cpp:765-778) followed by RayTracerBVH::setOctree's BFS flatten
 * (RayTracerBVH.cpp:443-490): voxels are x-fastest uint8 (0 EMPTY, 1 FILLED), dims in voxels.
 * Returns a malloc'ed array of *numNodes RtoGpuNode in the reference's BFS numbering; free with rto_host_free. */
RTO_API int rto_host_octree_build(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	RtoGpuNode** nodesOut, size_t* numNodes);

/* MarchingCubesRenderer::render(root, grid, 0,0,0, root->size) (Renderer.cpp:14-36 over localMC,
 * OctreeVoxel.cpp:780-879): triangle soup in the reference's emission order.  malloc'ed; rto_host_free. */
RTO_API int rto_host_mc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, const RtoGpuNode* nodes, size_t numNodes,
	RtoTriangle** trisOut, size_t* numTris);

/* The Adaptive Dual Contouring mesh (the triangle soup of configuration C4): renderOctree (main.cpp:95-208) calling
 * AdaptiveDualContouringRenderer::render -> createTriangles (AdaptiveDualContouringRenderer.cpp:489-803 with createFaceTriangles
 * :805-1088, gatherHermiteData :1090-1144, generateDualVertex :1146-1234, calculateIntersection :1236-1357, cellContainsSurface
 * :1367-1530, QEFSolver :46-160) on every leaf in depth-first order.  The renderer's dual-vertex cache makes the result depend on
 * that order; this builder keeps the order and gives the same triangles in the same sequence, bit for bit, with the expensive
 * per-cell work spread over all host threads.  viewProj16 == NULL: every leaf is visited; otherwise subtrees whose box, grown by
 * extraMargin (the reference passes 50), is outside the frustum of viewProj16 (rto_host_view_proj with zNear 0.01, zFar 5000) are
 * skipped like renderOctree does.  Octrees wider than 1024 voxels are refused (the reference's cell keys alias beyond 10 bits per
 * axis).  malloc'ed; rto_host_free. */
RTO_API int rto_host_dc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, const RtoGpuNode* nodes, size_t numNodes,
	const float* viewProj16 /* may be NULL */, float extraMargin, RtoTriangle** trisOut, size_t* numTris);
/* rto_host_dc_mesh that also returns the normal the reference stores with each triangle (MCTriangle::normal, the same vector three
 * times): normalize(cross(v1 - v0, v2 - v0)), negated when the emitting leaf is solid; the face normal for the fans of the boundary
 * fallback.  normalsOut: 3 floats per triangle, malloc'ed. */
RTO_API int rto_host_dc_mesh_normals(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, const RtoGpuNode* nodes, size_t numNodes,
	const float* viewProj16 /* may be NULL */, float extraMargin, RtoTriangle** trisOut, float** normalsOut, size_t* numTris);
/* The application's triangle cache (saveTriangleCache / loadTriangleCache, main.cpp:27-67; triangle_cache/dc_triangles_<hash>.bin):
 * size_t count, then count x MCTriangle { vec3 v[3]; vec3 normal[3]; } (72 bytes).  save: normals3 holds one normal per triangle
 * (written three times) or is NULL (the flat normal of the geometry, what localMC stores).  load: normals9Out (optional) receives
 * the three stored normals of every triangle.  malloc'ed; rto_host_free. */
RTO_API int rto_host_tricache_save(const char* path, const RtoTriangle* tris, const float* normals3 /* may be NULL */, size_t numTris);
RTO_API int rto_host_tricache_load(const char* path, RtoTriangle** trisOut, float** normals9Out /* may be NULL */, size_t* numTris);
/* The same mesh by replaying the reference's cache protocol leaf by leaf (the cross-check of rto_host_dc_mesh's order-free
 * formulation, and its way out if the fallback rounds did not settle). */
RTO_API int rto_host_dc_mesh_replay(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, const RtoGpuNode* nodes, size_t numNodes,
	const float* viewProj16 /* may be NULL */, float extraMargin, RtoTriangle** trisOut, size_t* numTris);

/* BVH::BVH(const std::vector<Triangle>&) (BVH.cpp:19-71): same tree (median split, longest axis, std::sort
 * by centroid, leaves <= 2 triangles).  The triangle array must outlive the handle (the reference keeps raw
 * pointers too, BVH.cpp:21-25). */
RTO_API int rto_host_bvh_build(const RtoTriangle* tris, size_t numTris, RtoHostBvh** out);
RTO_API void rto_host_bvh_free(RtoHostBvh* bvh);
RTO_API size_t rto_host_bvh_num_nodes(const RtoHostBvh* bvh);
/* Pre-order export (left before right): boxes6 = min xyz, max xyz; meta4 = isLeaf, triCount, tri0, tri1. */
RTO_API int rto_host_bvh_export(const RtoHostBvh* bvh, float* boxes6, int32_t* meta4, size_t capacity);

/* Camera(theta, phi, radius) with target (Camera.cpp:8-29), view = lookAt(eye, target, +Y), plus the two
 * host constants of generateRay: inverse(view) and tan(radians(fovDeg)/2).  Angles in radians. */
RTO_API int rto_host_camera_orbit(float theta, float phi, float radius, const float target[3],
	float fovDeg, float aspect, int width, int height, RtoCamera* out, float* view16 /* may be NULL */);

/* proj * view as RayTracerBVH::renderSceneComputeWithCulling forms it (RayTracerBVH.cpp:733-734): glm::perspective(radians(fovDeg),
 * aspect, zNear, zFar) times the view matrix of rto_host_camera_orbit; the reference passes zNear 0.01, zFar 5000. */
RTO_API int rto_host_view_proj(const float view16[16], float fovDeg, float aspect, float zNear, float zFar, float viewProj16[16]);

/* The CPU part of RayTracerBVH::renderSceneComputeWithCulling (RayTracerBVH.cpp:724-813 over Frustum.cpp:5-93), the variant the
 * reference's main loop actually calls (main.cpp:1357): every node whose world box, grown by `margin` (the reference passes 150),
 * lies entirely behind a frustum plane is dropped, the rest is compacted in index order and child indices are remapped (-1 for a
 * dropped child).  The result is the array the compute shader traverses; upload it with rto_scene_create_octree and render with
 * RTO_MODE_OCTREE_GLSL: hit ids are then indices into the CULLED array, newToOld (optional) maps them back.  malloc'ed; rto_host_free. */
RTO_API int rto_host_frustum_cull(const RtoGpuNode* nodes, size_t numNodes, const float gridMin[3], float voxelSize,
	const float viewProj16[16], float margin, RtoGpuNode** culledOut, size_t* numCulled, int32_t** newToOldOut /* may be NULL */);

/* loadVoxelGrid / saveVoxelGrid, CacheUtils.cpp:5-59 (sceneCache.bin: 3 x int32 dims, 4 x float, size_t n, bytes). */
RTO_API int rto_host_grid_load(const char* path, int dims[3], float minAndVoxel[4], uint8_t** voxelsOut);
RTO_API int rto_host_grid_save(const char* path, const int dims[3], const float minAndVoxel[4], const uint8_t* voxels);

/* loadCSVDataIntoVoxelGrid (BuildingLoader.cpp:153-290), the producer of sceneCache.bin: parses the DT vertex / face CSV files
 * (header line skipped; vertex lines "mesh,vertex,easting,northing,elevation,lat,lon,elevMin", face lines "mesh,v1,v2,v3"),
 * pads the bounds by one voxel, caps the grid at 1000 cells per axis by enlarging the voxel, and fills every voxel whose centre
 * projects into a face within the face's bounding box.  Same grid as the reference, bit for bit.  Empty / unreadable input gives
 * dims 0 and no voxels (the reference returns an empty VoxelGrid).  voxels: malloc'ed, x-fastest; rto_host_free. */
RTO_API int rto_host_csv_voxelize(const char* vertsCsv, const char* facesCsv, float voxelSize,
	int dims[3], float minAndVoxel[4], uint8_t** voxelsOut);

RTO_API void rto_host_free(void* p);

/* ---- device scenes ----------------------------------------------------------------------------------- */

/* Replaces RayTracerBVH::setOctree's SSBO upload (RayTracerBVH.cpp:492-504).  `nodes` is any GPUNodes array
 * with root at index 0 (normally rto_host_octree_build's).  The scene is linearised on upload into a
 * pointer-free array; arrays produced by the reference builder (every internal node has 8 contiguous
 * children) take the compact 4-byte-per-node path, anything else the general 64-byte-per-node path. */
RTO_API int rto_scene_create_octree(const RtoGpuNode* nodes, size_t numNodes, const float gridMin[3],
	float voxelSize, RtoScene** out);

/* Replaces "build a BVH, keep it for queries": builds the reference-shaped tree on the host (or takes `prebuilt`) and uploads
 * (a) that tree, for rto_bvh_query, rto_render_stats and RTO_FLAG_NO_PRUNE, and (b) the production tree the renders walk: a SAH
 * tree over single triangles whose records carry the box of their reference leaf, so that the candidate rule of BVH::query
 * ("every triangle of a leaf whose box passes intersectAABB") is applied exactly while the tree itself is free to be a good one
 * (DESIGN.md section 3).  Hit ids, t and colours equal the reference CPU path's bit for bit. */
RTO_API int rto_scene_create_bvh(const RtoTriangle* tris, size_t numTris, const RtoHostBvh* prebuilt /* may be NULL */,
	RtoScene** out);

/* ---- scene construction on the GPU (the steps in front of the ray path) ----------------------------
 * Same results as the host builders above, bit for bit (the tests compare both), at memory-bandwidth speed. */

/* createOctreeFromVoxelGrid + setOctree numbering on the device: same array as rto_host_octree_build (malloc'ed; rto_host_free). */
RTO_API int rto_device_octree_build(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	RtoGpuNode** nodesOut, size_t* numNodes);

/* Voxel grid -> device octree scene without a host-side tree: upload the grid, build and linearise on the GPU
 * (replaces createOctreeFromVoxelGrid + RayTracerBVH::setOctree, main.cpp:1077, 1127-1131). */
RTO_API int rto_scene_create_octree_from_grid(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, RtoScene** out);

/* MarchingCubesRenderer::render(root, grid, 0,0,0, root->size) on the device: same triangle soup, same order as
 * rto_host_mc_mesh (malloc'ed; rto_host_free).  Builds the octree it needs for the emission order itself. */
RTO_API int rto_device_mc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, RtoTriangle** trisOut, size_t* numTris);

/* rto_host_csv_voxelize with the fill on the GPU (one warp per face); same grid, bit for bit. */
RTO_API int rto_device_csv_voxelize(const char* vertsCsv, const char* facesCsv, float voxelSize,
	int dims[3], float minAndVoxel[4], uint8_t** voxelsOut);

/* The Adaptive Dual Contouring mesh extracted on the GPU: same arguments and same result as rto_host_dc_mesh (same triangles, same
 * order, bit for bit).  The reference's visit-order-dependent vertex cache is resolved without a sequential pass: every cache key
 * takes the vertex of its first toucher, found as an atomic minimum over (Morton position of the visiting leaf, kind of touch),
 * and the boundary fallback (createFaceTriangles) is settled in rounds (csrc/rto_dc.cu).  malloc'ed; rto_host_free. */
RTO_API int rto_device_dc_mesh(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, const RtoGpuNode* nodes, size_t numNodes,
	const float* viewProj16 /* may be NULL */, float extraMargin, RtoTriangle** trisOut, size_t* numTris);

/* rto_host_frustum_cull on the GPU (test, prefix sum, compaction + remap); same arrays. */
RTO_API int rto_device_frustum_cull(const RtoGpuNode* nodes, size_t numNodes, const float gridMin[3], float voxelSize,
	const float viewProj16[16], float margin, RtoGpuNode** culledOut, size_t* numCulled, int32_t** newToOldOut /* may be NULL */);

/* BVH scene built on the device: the fast, NOT reference-shaped route (linear BVH: Morton sort, leaves of two neighbours, radix
 * tree, exact union boxes).  BVH::build's tree depends on std::sort's order on equal centroids (BVH.cpp:58-60) and cannot be
 * rebuilt in parallel; rto_scene_create_bvh stays the route whose hit ids equal the reference's exactly.  Here hit ids (indices
 * into `tris`) can differ only at near-ties (equal t on a shared edge, a ray grazing a leaf-box edge).  rto_bvh_query and
 * rto_render_stats need the reference tree and return RTO_ERR_UNSUPPORTED on such a scene. */
RTO_API int rto_scene_create_bvh_device(const RtoTriangle* tris, size_t numTris, RtoScene** out);

/* Voxel grid -> octree -> Marching-Cubes mesh -> linear BVH, all on the GPU, nothing but the grid crosses the bus
 * (replaces createOctreeFromVoxelGrid + MarchingCubesRenderer::render + BVH::BVH for the mesh path).  Hit ids index the
 * triangle soup rto_host_mc_mesh / rto_device_mc_mesh return for the same grid. */
RTO_API int rto_scene_create_bvh_from_grid(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, RtoScene** out);
/* The same with the Adaptive Dual Contouring mesher in the middle (the pipeline configuration C4 names): grid -> octree -> DC mesh
 * (rto_device_dc_mesh's triangles and order; viewProj16 / extraMargin as there) -> linear BVH, nothing leaves the device. */
RTO_API int rto_scene_create_bvh_from_grid_dc(const uint8_t* voxels, int dimX, int dimY, int dimZ,
	const float gridMin[3], float voxelSize, const float* viewProj16 /* may be NULL */, float extraMargin, RtoScene** out);

/* Diagnostic: the compact device layout of an octree scene (desc: numNodes + 8 words, up: (numNodes + 7) / 8 + 1 words,
 * inner: 4 words per internal node); any pointer may be NULL.  Lets tests compare the two construction routes. */
RTO_API int rto_scene_octree_layout_read(RtoScene* scene, uint32_t* desc, int32_t* up, int32_t* inner4, size_t* numInner);

/* A device scene as one file, and back: the flattened arrays the kernels walk (BVH: production and reference-shaped nodes, triangle
 * records, wide form; octree: compact or general arrays), whatever route built them.  The reference rebuilds octree, flatten and BVH at
 * every start (main.cpp:1127-1131) and caches only what comes before them -- sceneCache.bin (CacheUtils.cpp:5-59, rto_host_grid_*) and the
 * Dual-Contouring triangle cache (main.cpp:27-92, rto_host_tricache_*); this is the cache of what comes after.
 * rto_scene_load gives a scene that renders the same bits as the one saved.  A file written by another version of the library is refused
 * with RTO_ERR_UNSUPPORTED, a truncated or damaged one (checksum) with RTO_ERR_IO. */
RTO_API int rto_scene_save(RtoScene* scene, const char* path);
RTO_API int rto_scene_load(const char* path, RtoScene** out);

RTO_API void rto_scene_destroy(RtoScene* scene);
RTO_API int rto_scene_info(const RtoScene* scene, int* kind /* RtoMode of a BVH or octree scene */,
	size_t* numPrims, size_t* numNodes, size_t* deviceBytes, int* compactLayout);
/* The CUDA stream all work of this scene is enqueued on (cudaStream_t as void*). */
RTO_API void* rto_scene_stream(const RtoScene* scene);

/* ---- rendering (replaces RayTracerBVH::renderSceneCompute, RayTracerBVH.cpp:614-704) -------------------- */

/* Trace rows [y0, y1) of the camera's image.  Planes in `frame` are indexed from row y0 (size (y1-y0)*width).
 * mode must match the scene kind (BVH scene: RTO_MODE_BVH; octree scene: either octree mode).
 * shadowBias: offset of the shadow-ray origin along the shading normal (rule: 1e-3f * scene scale).
 * With RTO_MEM_HOST the call copies results to the host and synchronises -- a frame is traced in four row bands (a batch frame by
 * frame), each band's copy running while the next is traced; the copies reach the speed of the link only into page-locked planes
 * (rto_host_alloc_pinned) --; with RTO_MEM_DEVICE it only enqueues on rto_scene_stream() (use rto_scene_sync or your own event). */
RTO_API int rto_render(RtoScene* scene, const RtoCamera* cam, int mode, uint32_t flags, float shadowBias,
	int y0, int y1, const RtoFrame* frame);

/* Many cameras in one submission (frame f writes planes at offset f*(y1-y0)*width). */
RTO_API int rto_render_batch(RtoScene* scene, const RtoCamera* cams, int numCams, int mode, uint32_t flags,
	float shadowBias, int y0, int y1, const RtoFrame* frame);

/* ---- compact frames: 4 bytes per pixel out of the trace kernel ---------------------------------------------------------------------
 * Colour, hit id and t of a pixel of a BVH scene are pure functions of (camera, pixel, hit triangle, shadow bit).  rto_render_codes
 * writes only that: one 32-bit word per pixel (0 = miss, else (position of the hit triangle in the scene's leaf order + 1) |
 * normal flipped << 30 | shadowed << 31), in the kernel's TILE order (the 128 pixels of a 16 x 8 block are consecutive: a warp writes one full 128-byte line,
 * which is what makes the destination usable across NVLink), and rto_resolve_codes rebuilds the three planes of RtoFrame from the
 * words with the same ray generation, the same Moller-Trumbore arithmetic on the same triangle record and the same shading: the planes
 * equal those of rto_render_batch bit for bit (tests/test_gpu_codes.py).  The words are meaningful to any scene built from the same
 * triangles by the same route (scenes replicated over the GPUs of a box), not across different scenes.
 * This is the path's one exchange step (the reference keeps its frame in a GL texture on the one GPU it has, RayTracerBVH.cpp:815-887):
 * a GPU that traces frames for another one ships 4 bytes per pixel instead of 24.
 * y0 must be a multiple of 8 and y1 a multiple of 8 or the image height.  `codes` addresses frame 0 of the code buffer; the call fills
 * frames [firstFrame, firstFrame + numCams), rto_codes_frame_words(width, height) words each.  `stream` is a cudaStream_t of the scene's
 * device, or NULL for the scene's own stream.  Both calls only enqueue (rto_resolve_codes with RTO_MEM_HOST also copies and waits). */
RTO_API size_t rto_codes_frame_words(int width, int height);
RTO_API int rto_render_codes(RtoScene* scene, const RtoCamera* cams, int numCams, uint32_t flags, float shadowBias, int y0, int y1,
	uint32_t* codes /* device-accessible: local, peer or IPC-mapped memory */, size_t firstFrame, void* stream);
RTO_API int rto_resolve_codes(RtoScene* scene, const RtoCamera* cams, int numCams, int y0, int y1, const uint32_t* codes, size_t firstFrame,
	const RtoFrame* frame /* planes indexed from (frame 0 of this call, row y0) */, void* stream);

/* Exchange memory: a zero-filled device buffer GPUs of the same box write hit codes into.  Same process: allocate on the receiving device and let
 * the senders enable peer access (rto_group_* does).  Other processes (one process per GPU): pass the 64-byte handle over any channel and
 * map it with rto_exchange_open (CUDA IPC; the mapping is an ordinary device pointer for rto_render_codes). */
typedef struct RtoIpcHandle { unsigned char bytes[64]; } RtoIpcHandle;
RTO_API int rto_exchange_alloc(size_t bytes, void** devPtr, RtoIpcHandle* handleOut /* may be NULL */);
RTO_API int rto_exchange_free(void* devPtr);
RTO_API int rto_exchange_open(const RtoIpcHandle* handle, void** devPtr);
RTO_API int rto_exchange_close(void* devPtr);

/* ---- device groups: several GPUs of one box behind one handle (one process, one host thread) ----------------------------------------
 * What the reference's main loop would call instead of one renderSceneCompute per frame (main.cpp:1357-1363) when the box has more
 * than one GPU: the scene is replicated on every device (trees built once on the host), the rows of a batch of frames are dealt to the
 * devices in contiguous ranges, every device but the first writes hit codes straight into the first device's memory from its trace
 * kernel (NVLink peer stores), and the first device expands them into the caller's planes on a second stream while it traces its own
 * share.  Shares follow the measured time of every device.  The planes equal rto_render_batch's bit for bit.
 * frame->memory: RTO_MEM_DEVICE = memory of devices[0] (the call only enqueues; order your own work after rto_group_stream() or call
 * rto_group_sync), RTO_MEM_HOST = copied and waited for. */
typedef struct RtoGroup RtoGroup;
RTO_API int rto_group_create(const int* devices, int numDevices, RtoGroup** out);
RTO_API void rto_group_destroy(RtoGroup* group);
RTO_API int rto_group_size(const RtoGroup* group);
/* rto_scene_create_bvh on every device of the group (`prebuilt` as there). */
RTO_API int rto_group_scene_bvh(RtoGroup* group, const RtoTriangle* tris, size_t numTris, const RtoHostBvh* prebuilt /* may be NULL */);
/* The replica on devices[member] (owned by the group). */
RTO_API RtoScene* rto_group_scene(RtoGroup* group, int member);
RTO_API int rto_group_render_batch(RtoGroup* group, const RtoCamera* cams, int numCams, uint32_t flags, float shadowBias, const RtoFrame* frame);
/* enabled: adapt the shares to the measured times (default on); weights (optional, one per device): shares to start from or to keep;
 * chunks: launches per device and batch, so that the first device expands chunk c while chunk c + 1 is traced (default 4; 0 = keep). */
RTO_API int rto_group_set_balancing(RtoGroup* group, int enabled, const float* weights /* may be NULL */, int chunks);
/* Device time of the last batch on every device (the first device: its own trace, with the expansion running beside it).  Waits. */
RTO_API int rto_group_last_ms(RtoGroup* group, float* msPerDevice);
RTO_API void* rto_group_stream(const RtoGroup* group);
RTO_API int rto_group_sync(RtoGroup* group);

/* Explicit ray list (origins/directions 3 floats each, directions need not be unit), for edge cases and as the
 * batched counterpart of single-ray calls: octree scenes return what octreeRaySkip(root, ro, rd, tMin, tMax, grid)
 * returns (mode A) or the GLSL traversal's closestT (mode B); BVH scenes the closest MT hit.  [tMin, tMax] is octreeRaySkip's
 * interval; the GLSL traversal and the BVH hit rule have none: BVH scenes accept only tMin = 0, tMax >= 1e30 (RTO_ERR_INVALID
 * otherwise).  memory: RTO_MEM_HOST or RTO_MEM_DEVICE for all four arrays. */
RTO_API int rto_trace_rays(RtoScene* scene, int mode, uint32_t flags, const float* origins, const float* dirs, size_t numRays,
	float tMin, float tMax, float* tOut, int32_t* idOut, int memory);

/* Page-locked host memory for frame planes (cudaHostAlloc, usable from every device): RTO_MEM_HOST frames are copied at the speed of the
 * link only into page-locked memory; into ordinary memory the driver stages the copy (measured: profiles/README.md).  The reference never
 * reads its frame back (a GL texture, RayTracerBVH.cpp:815-887), so there is no counterpart.  RTO_ERR_NO_DEVICE without a CUDA device. */
RTO_API int rto_host_alloc_pinned(size_t bytes, void** out);
RTO_API void rto_host_free_pinned(void* p);

/* The stable device radix sort behind RTO_FLAG_SORT_RAYS on its own (csrc/rto_sort.cuh; hand-written, no library sort): n pairs of 32-bit
 * keys and values in host memory, sorted by key in place.  Not in the reference (north-star extension: ray coherence sorting). */
RTO_API int rto_device_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n);

/* BVH::query for many rays (BVH.cpp:107-113): candidate triangle ids per ray in the reference's order.
 * offsets has numRays+1 entries (host memory).  Call with ids == NULL to get the total in *totalOut first. */
RTO_API int rto_bvh_query(RtoScene* scene, const float* origins, const float* dirs, size_t numRays,
	int64_t* offsets, int32_t* ids, size_t idsCapacity, size_t* totalOut);

/* Work counters of the reference algorithm for the image rows [y0,y1), measured on the GPU by replaying the
 * reference's visit pattern: BVH scenes: {box tests primary, candidates primary, box tests shadow, candidates shadow,
 * shadow rays}; octree scenes: {nodes visited, 0, 0, 0, 0}.  Used for the algorithmic bytes/flops of SURVEY.md 8d. */
RTO_API int rto_render_stats(RtoScene* scene, const RtoCamera* cam, int mode, uint32_t flags, float shadowBias,
	int y0, int y1, uint64_t stats[5]);

/* VolumeRaycastRenderer's per-frame use of octreeRaySkip (VolumeRaycastRenderer.cpp:1598-1664): 49 probe rays (7 x 7 over the central
 * +-0.2 of NDC, unprojected through inverse(perspective(45 deg, aspect, 0.1, 5000)) and inverse(view)) are traced with
 * octreeRaySkip semantics on the GPU; the 15th percentile of the valid results, times 0.75, blended 0.4 : 0.6 with
 * lastSkipDistance (the reference keeps that in a function-local static) is the distance the volume ray marcher may skip.
 * probeT (optional) receives the 49 raw results. */
RTO_API int rto_octree_skip_distance(RtoScene* scene, const float view16[16], const float camPos[3], float aspect,
	float lastSkipDistance, float* skipDistanceOut, float* probeT /* may be NULL */);
/* The two host halves of it: the 49 rays (origins / dirs: 147 floats each) and the percentile + blend. */
RTO_API int rto_host_skip_probe_rays(const float view16[16], const float camPos[3], float aspect, float* origins, float* dirs);
RTO_API float rto_host_skip_distance_from_probes(const float* t, int count, float lastSkipDistance);

RTO_API int rto_scene_sync(RtoScene* scene);
/* Device time in milliseconds of the most recent rto_render / rto_render_batch on this scene
 * (CUDA events on the scene's stream around the kernels only).  Synchronises. */
RTO_API int rto_scene_last_kernel_ms(RtoScene* scene, float* ms);
/* Number of kernel launches issued by this library on this scene so far. */
RTO_API uint64_t rto_scene_launch_count(const RtoScene* scene);

#ifdef __cplusplus
}
#endif
#endif /* RTO_C_H */
