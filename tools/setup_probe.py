"""where the host-side set-up time of a large mesh goes (A2DS_VERBOSE=1 prints the stages of
the library; this prints the stages of the Python face).  python tools/setup_probe.py [nt nx]"""
import importlib, sys, time
import numpy as np
sys.path.insert(0, ".")
a2ds = importlib.import_module("a2d-shells_b200")
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
T = [time.time()]
def lap(what):
    T.append(time.time()); print(f"[probe] {what:28s} {T[-1] - T[-2]:.3f} s", flush=True)
conn, X, ends = a2ds.meshes.cylinder(nt, nx); lap("mesh arrays (numpy)")
u = a2ds.meshes.seeded_state(np.arange(len(X)), 1e-5); lap("seeded state (numpy)")
Cs, eth = a2ds.iso_shell_tables()
asm = a2ds.Assembler(0); lap("a2ds_create")
asm.set_mesh(conn, len(X)); lap("set_mesh")
asm.set_nodes(X); asm.set_components(Cs[None], eth[None]); asm.set_state(u); asm.synchronize(); lap("nodes, components, state")
k = asm.create_mat(); lap("create_mat #1")
g = asm.create_mat(); lap("create_mat #2")
asm.assembleAll(k, g, download=False); asm.synchronize(); lap("first assembly (lists, launch)")
asm.assembleAll(k, g, download=False); asm.synchronize(); lap("second assembly")
asm.close(); lap("close")
