"""Byte-level legacy-VTK fixtures in the layouts VTK's own writers emit (TEST INFRASTRUCTURE).

pyvista / vtk are not installed here, so these files are NOT produced by them; they are written by this
independent script following the VTK file-format specification ("VTK File Formats", legacy section) and the
conventions of vtkDataWriter that the reference's files carry (``mesh.save("*.vtk")``, generate_dataset.py:584-587):

  * header ``# vtk DataFile Version 5.1`` (VTK >= 9.1) or ``4.2`` / ``3.0`` (older), title ``vtk output``;
  * ASCII arrays: values separated by one blank, 9 per line, every line ENDING with a blank, the last line too;
  * BINARY arrays: big-endian raw values followed by one ``\\n``;
  * 5.1 cell layout: ``CELLS <n_cells + 1> <n_connectivity>`` / ``OFFSETS vtktypeint64`` / ``CONNECTIVITY vtktypeint64``;
    classic layout: ``CELLS <n_cells> <n_cells + n_connectivity>`` then ``k i0 .. ik-1`` per cell as ``int``;
  * ``CELL_TYPES n`` with one type per line (5 = VTK_TRIANGLE, 9 = VTK_QUAD);
  * an optional ``METADATA`` / ``INFORMATION`` block after an array, closed by an empty line;
  * ``POINT_DATA`` / ``CELL_DATA`` / ``FIELD`` attribute sections after the geometry.

    python tests/golden/make_vtk_fixtures.py      # rewrites tests/golden/vtk/*.vtk

This script shares no code with ``pdivgnn_b200.io`` (reader under test).
"""
import os
import struct

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vtk")

# a 3 x 2 patch: 6 points, 4 triangles / 2 quads
POINTS = [(0.0, 0.0, 0.0), (1.5, 0.0, 0.0), (3.0, 0.0, 0.0), (0.0, 1.0, 0.0), (1.5, 1.0, 0.0), (3.0, 1.0, 0.0)]
TRIS = [(0, 1, 4), (0, 4, 3), (1, 2, 5), (1, 5, 4)]
QUADS = [(0, 1, 4, 3), (1, 2, 5, 4)]


def ascii_array(vals, fmt):
    out = []
    for i in range(0, len(vals), 9):
        out.append("".join((fmt % v) + " " for v in vals[i:i + 9]) + "\n")
    return "".join(out).encode()


def be(vals, code):
    return struct.pack(">%d%s" % (len(vals), code), *vals) + b"\n"


METADATA = (b"METADATA\nINFORMATION 2\nNAME L2_NORM_RANGE LOCATION vtkDataArray\nDATA 2 0 3.16228 \n"
            b"NAME L2_NORM_FINITE_RANGE LOCATION vtkDataArray\nDATA 2 0 3.16228 \n\n")


def write(name, version, binary, dataset, cells, with_metadata, with_attributes):
    k = len(cells[0])
    flat_pts = [c for p in POINTS for c in p]
    conn = [i for c in cells for i in c]
    b = bytearray()
    b += ("# vtk DataFile Version %s\nvtk output\n%s\nDATASET %s\n" % (version, "BINARY" if binary else "ASCII", dataset)).encode()
    b += ("POINTS %d double\n" % len(POINTS)).encode()
    b += be(flat_pts, "d") if binary else ascii_array(flat_pts, "%g")
    if with_metadata:
        b += METADATA
    sec = "POLYGONS" if dataset == "POLYDATA" else "CELLS"
    if version.startswith("5"):
        offs = list(range(0, k * len(cells) + 1, k))
        b += ("%s %d %d\nOFFSETS vtktypeint64\n" % (sec, len(cells) + 1, len(conn))).encode()
        b += be(offs, "q") if binary else ascii_array(offs, "%d")
        b += b"CONNECTIVITY vtktypeint64\n"
        b += be(conn, "q") if binary else ascii_array(conn, "%d")
    else:
        stream = [v for c in cells for v in (k,) + tuple(c)]
        b += ("%s %d %d\n" % (sec, len(cells), len(stream))).encode()
        if binary:
            b += be(stream, "i")
        else:  # classic ASCII: one cell per line
            b += "".join("%d %s \n" % (k, " ".join(str(i) for i in c)) for c in cells).encode()
    if dataset == "UNSTRUCTURED_GRID":
        types = [5 if k == 3 else 9] * len(cells)
        b += ("CELL_TYPES %d\n" % len(cells)).encode()
        b += be(types, "i") if binary else "".join("%d\n" % t for t in types).encode()
    if with_attributes:
        b += ("\nPOINT_DATA %d\nSCALARS Stress float\nLOOKUP_TABLE default\n" % len(POINTS)).encode()
        vals = [0.5 * i for i in range(len(POINTS))]
        b += be(vals, "f") if binary else ascii_array(vals, "%g")
    with open(os.path.join(HERE, name), "wb") as fh:
        fh.write(bytes(b))


def main():
    os.makedirs(HERE, exist_ok=True)
    write("tri_ug_v51_ascii.vtk", "5.1", False, "UNSTRUCTURED_GRID", TRIS, True, True)
    write("tri_ug_v51_binary.vtk", "5.1", True, "UNSTRUCTURED_GRID", TRIS, True, True)
    write("tri_ug_v42_ascii.vtk", "4.2", False, "UNSTRUCTURED_GRID", TRIS, False, True)
    write("tri_pd_v51_ascii.vtk", "5.1", False, "POLYDATA", TRIS, False, False)
    write("quad_ug_v42_binary.vtk", "4.2", True, "UNSTRUCTURED_GRID", QUADS, False, False)
    write("quad_ug_v51_ascii.vtk", "5.1", False, "UNSTRUCTURED_GRID", QUADS, True, False)
    write("quad_pd_v30_ascii.vtk", "3.0", False, "POLYDATA", QUADS, False, True)


if __name__ == "__main__":
    main()
