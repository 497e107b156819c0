"""9-node shells (TACSQuad9Shell) on one GPU: residual + tangent of an n x n plate of 9-node
elements (same node count as a 2n x 2n plate of 4-node elements), device resident.
python tools/quad9_bench.py [n=500] [steps=5] -> one JSON line"""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
a2ds = importlib.import_module("a2d-shells_b200")


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    conn, X, bcn = a2ds.meshes.plate9(n, n, bump=1e-3)
    nn = len(X)
    Cs, eth = a2ds.iso_shell_tables()
    asm = a2ds.Assembler(0)
    asm.set_mesh(conn, nn, order=3)
    asm.set_nodes(X)
    asm.set_components(Cs[None], eth[None])
    asm.set_bcs(bcn, 63)
    asm.set_state(a2ds.meshes.seeded_state(np.arange(nn), 1e-5))
    kmat = asm.create_mat()
    for _ in range(3):
        asm.assembleJacobian(1.0, 0.0, 0.0, kmat, download=False)
    asm.synchronize()
    asm.region_begin()
    for _ in range(steps):
        asm.assembleJacobian(1.0, 0.0, 0.0, kmat, download=False)
    ms = asm.region_end() / steps
    kms = asm.last_kernel_ms()
    # fused residual + tangent + geometric stiffness (the buckling flow's three outputs)
    gmat = asm.create_mat()
    for _ in range(2):
        asm.assembleAll(kmat, gmat, download=False)
    asm.synchronize()
    asm.region_begin()
    for _ in range(steps):
        asm.assembleAll(kmat, gmat, download=False)
    ms_all = asm.region_end() / steps
    out = {"workload": f"plate {n}x{n} 9-node MITC shells (TACSQuad9Shell), residual + tangent into BCSR6",
           "elements": len(conn), "nodes": nn, "blocks": int(asm.mat_nnz(kmat)),
           "ms_per_step": ms, "kernel_ms": kms, "elements_per_s": len(conn) / (ms * 1e-3),
           "nodes_per_s": nn / (ms * 1e-3),
           "res_K_G": {"ms_per_step": ms_all, "elements_per_s": len(conn) / (ms_all * 1e-3)}}
    # sampled parity against the order-3 oracle on the elements around a few nodes (outside the timing)
    try:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle_py as orc
        res = asm.assembleJacobian(1.0, 0.0, 0.0, kmat)
        rng = np.random.default_rng(1)
        sample = rng.choice(len(conn), size=64, replace=False)
        # rows of nodes interior to an element (its centre node 4) get contributions of that element only
        comp = orc.make_comp(0, Cs, eth)
        u = a2ds.meshes.seeded_state(np.arange(nn), 1e-5)
        worst = 0.0
        bc = set(int(b) for b in bcn)
        for e in sample:
            r_o, _ = orc.jacobian(comp, X[conn[e]].ravel(), u[conn[e]].ravel(), order=3)
            c = int(conn[e, 4])
            if c in bc:
                continue
            worst = max(worst, np.abs(res[c] - r_o[24:30]).max() / np.abs(r_o).max())
        out["parity_res_centre_nodes_max_rel"] = worst
    except Exception as ex:  # noqa: BLE001
        out["parity_error"] = str(ex)
    asm.close()
    # the unmodified reference beside it (oracle/_ref): assembleJacobian of TACSQuad9Shell on a
    # bounded sample of the same mesh family, threaded as the reference allows
    try:
        import refdrv
        if refdrv.available():
            nr = min(n, 120)
            conn_r, X_r, bcn_r = a2ds.meshes.plate9(nr, nr, bump=1e-3)
            ra = refdrv.RefAssembler(conn_r, X_r, np.zeros(len(conn_r), dtype=np.int32),
                                     refdrv.iso_props(kind=2)[None], bcn_r, [list(range(6))] * len(bcn_r),
                                     [[0.0] * 6] * len(bcn_r), nodes_per_elem=9)
            ur = np.zeros((len(X_r), 6))
            ur[ra.new_nodes] = a2ds.meshes.seeded_state(np.arange(len(X_r)), 1e-5)
            ra.set_state(ur)
            km = ra.mat_create(2)
            threads = min(16, os.cpu_count() or 1)
            ra.set_threads(threads)
            ra.time(1, km)
            t = float(np.median([ra.time(1, km) for _ in range(3)]))
            ra.close()
            out["cpu_reference"] = {"elements_per_s": len(conn_r) / t, "cores": threads, "kind": "reference",
                                    "sample": f"plate {nr}x{nr} 9-node elements, assembleJacobian into TACSSchurMat, "
                                              f"median of 3, {threads} pthreads"}
    except Exception as ex:  # noqa: BLE001
        out["cpu_reference_error"] = str(ex)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
