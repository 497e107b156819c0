"""Host-side timing of the mesh input path (no GPU): the reference's TACSMeshLoader
(oracle/_ref, test infrastructure) beside a2ds_mesh_read_bdf with 1 and N threads and the
binary container, on a generated plate deck.   python tools/bdf_bench.py [nx] [fmt]"""
import importlib
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
a2ds = importlib.import_module("a2d-shells_b200")


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    fmt = sys.argv[2] if len(sys.argv) > 2 else "large"
    conn, X, bcn = a2ds.meshes.plate(nx, nx)
    d = tempfile.mkdtemp()
    deck, binf = os.path.join(d, "plate.bdf"), os.path.join(d, "plate.a2dm")
    t0 = time.time()
    a2ds.meshes.write_bdf(deck, conn, X, bcn, ["123456"] * len(bcn), 0.0, None, fmt)
    mb = os.path.getsize(deck) / 1e6
    print(f"deck: plate {nx}x{nx}, {len(conn)} CQUAD4, {len(X)} GRID, format {fmt}, {mb:.1f} MB "
          f"(written in {time.time() - t0:.1f} s), host cores: {os.cpu_count()}")
    rows = []
    try:
        import refdrv
        if refdrv.available():
            fail, ref, secs = refdrv.bdf_scan(deck)
            assert fail == 0
            rows.append(("reference TACSMeshLoader::scanBDFFile (1 thread)", secs))
        else:
            ref = None
    except Exception as e:  # the comparison is optional
        print("reference loader unavailable:", e)
        ref = None
    for nt in (1, 2, 4, 8, 16):
        if nt > 2 * (os.cpu_count() or 1):
            break
        t0 = time.time()
        m = a2ds.Mesh.L_read(deck, nt) if hasattr(a2ds.Mesh, "L_read") else None
        # time only the C call, not the copies into numpy arrays
        import ctypes as C
        L = a2ds.load_library()
        h = C.c_void_p()
        t0 = time.time()
        rc = L.a2ds_mesh_read_bdf(C.c_char_p(deck.encode()), C.c_int(nt), C.byref(h))
        secs = time.time() - t0
        assert rc == 0
        rows.append((f"a2ds_mesh_read_bdf, {nt} thread(s)", secs))
        if nt == 1:
            m = a2ds.Mesh(h)
            if ref is not None:
                same = all(np.array_equal(getattr(m, k), ref[k]) for k in
                           ("elem_ptr", "elem_conn", "elem_comp", "bc_nodes", "bc_ptr", "bc_vars"))
                same = same and m.X.tobytes() == ref["X"].tobytes()
                print("arrays identical to the reference loader's:", same)
            t0 = time.time(); m.write_bin(binf); tw = time.time() - t0
        else:
            L.a2ds_mesh_free(h)
    h = C.c_void_p()
    t0 = time.time()
    assert L.a2ds_mesh_read_bin(C.c_char_p(binf.encode()), C.byref(h)) == 0
    rows.append((f"a2ds_mesh_read_bin ({os.path.getsize(binf) / 1e6:.1f} MB container, written in {tw:.2f} s)",
                 time.time() - t0))
    L.a2ds_mesh_free(h)
    for name, secs in rows:
        print(f"  {name:78s} {secs:8.3f} s  {mb / secs:8.1f} MB/s  {len(conn) / secs / 1e6:7.2f} M elem/s")
    os.remove(deck); os.remove(binf); os.rmdir(d)


if __name__ == "__main__":
    main()
