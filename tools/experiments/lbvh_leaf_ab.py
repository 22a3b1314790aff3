import os,sys,time
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import ray_tracing_octrees_b200 as rto
assert rto.lib().rto_init(0)==0
W,H=1920,1080
for name in ("c1","dt","c3"):
    g = rto.generate_test_volume(128) if name=="c1" else (rto.VoxelGrid.load('/root/repo/tests/golden/dt_sceneCache.bin.gz') if name=="dt" else rto.city_block_grid(512,1234,32))
    ext=float(max(g.dims)*g.voxel_size)
    cams=[rto.Camera.from_degrees(35,40.0+45.0*k,(0.6 if name=="dt" else 0.9)*ext).consts(45.0,float(np.float32(W)/np.float32(H)),W,H)[0] for k in range(4)]
    sc=rto.Scene.bvh_from_grid(g) if os.environ.get("AB_ROUTE","device")=="device" else rto.Scene.bvh(rto.marching_cubes_mesh(g, rto.create_octree_from_voxel_grid(g)))
    F=4
    rgba=torch.empty((F,H,W,4),dtype=torch.float32,device="cuda"); hid=torch.empty((F,H,W),dtype=torch.int32,device="cuda"); tt=torch.empty((F,H,W),dtype=torch.float32,device="cuda")
    out=[]
    for flags in (rto.FLAG_SHADOWS, 0):
        ms=[]
        for _ in range(4):
            sc.render_device(cams,rto.MODE_BVH,flags,1e-3*g.voxel_size,0,H,rgba.data_ptr(),hid.data_ptr(),tt.data_ptr()); ms.append(sc.last_kernel_ms())
        out.append("%.3f"%np.median(ms[1:]))
    print(name, os.environ.get("AB_ROUTE","device"), "grow 2^-%s"%os.environ.get("RTO_BVH_GROW_LOG2","16"), "render ms shadows/primary", out, "idsum", int(hid.to(torch.int64).sum().item()))
