"""Synthetic DT-style CSV pairs for the voxeliser tests (the reference checkout lacks DT/DTVerts.csv and DT/DTFaces.csv, see
.MISSING_LARGE_BLOBS): box "buildings" made of triangles, with the irregularities the reference's loaders tolerate."""
import numpy as np

VERT_HEADER = "Mesh Number,Vertex Number,Easting,Northing,Elevation,Latitude,Longitude,Elev Min\n"
FACE_HEADER = "Mesh Number,Vertex 1,Vertex 2,Vertex 3\n"


def write_city_csv(tmpdir, seed=0, buildings=12, extent=400.0, messy=True, origin=(-5000.0, 5650000.0, 1040.0)):
    """-> (verts_path, faces_path).  Coordinates look like UTM eastings/northings (large offsets), which is what the float casts of
    the reference see."""
    rng = np.random.default_rng(seed)
    vl, fl = [VERT_HEADER], [FACE_HEADER]
    for m in range(buildings):
        cx, cy = rng.uniform(0, extent, 2)
        w, d, h = rng.uniform(8, 40), rng.uniform(8, 40), rng.uniform(5, 120)
        rot = rng.uniform(0, np.pi)
        corners = []
        for dz in (0.0, h):
            for sx, sy in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                x, y = sx * w / 2, sy * d / 2
                corners.append((origin[0] + cx + x * np.cos(rot) - y * np.sin(rot), origin[1] + cy + x * np.sin(rot) + y * np.cos(rot), origin[2] + dz))
        for k, (x, y, z) in enumerate(corners):
            vl.append("%d,%d,%.6f,%.6f,%.6f,51.0,-114.0,%.3f\n" % (m, k, x, y, z, origin[2]))
        quads = [(0, 1, 2, 3), (4, 5, 6, 7), (0, 1, 5, 4), (1, 2, 6, 5), (2, 3, 7, 6), (3, 0, 4, 7)]
        for a, b, c, d4 in quads:
            fl.append("%d,%d,%d,%d\n" % (m, a, b, c))
            fl.append("%d, %d ,%d,%d\n" % (m, a, c, d4))          # blanks around a field are trimmed
    if messy:
        vl.insert(5, "\n")                                          # empty line
        vl.insert(9, "3,notanumber,1,2,3,4,5,6\n")                  # stoi throws: line skipped
        vl.insert(12, "1,2,3\n")                                    # too few fields
        vl.append("0,0,%.6f,%.6f,%.6f,51.0,-114.0,0\n" % (origin[0] + 11.0, origin[1] + 13.0, origin[2] + 2.0))   # redefinition: last one wins
        vl.append("2,1,nan,1,1,0,0,0\n")                            # parses, not finite: kept in the map, ignored by the bounds
        fl.append("99,0,1,2\n")                                     # unknown mesh
        fl.append("1,0,1,77\n")                                     # unknown vertex
        fl.append("1,0,1\n")                                        # too few fields
        fl.append("1,0x,1,2\n")                                     # stoi parses the leading "0"
        fl.append("4,0,0,0\n")                                      # degenerate face
    vp, fp = str(tmpdir / "verts.csv"), str(tmpdir / "faces.csv")
    open(vp, "w").write("".join(vl))
    open(fp, "w").write("".join(fl))
    return vp, fp
