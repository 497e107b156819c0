"""The link-time drop-in: the UNMODIFIED reference buckling flow with
libtacs_a2ds_shim.so preloaded (its TACSAssembler::assembleRes/Jacobian/MatType run on the
GPU) against the same flow without it."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "tests", "_shim", "libtacs_a2ds_shim.so")


def _run(preload, *args):
    env = dict(os.environ)
    env["OPENBLAS_NUM_THREADS"] = "1"
    if preload:
        # the shim pulls in the reference library, whose bundled OpenBLAS needs its sibling
        # libgfortran on the loader path
        import sysconfig
        blas = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs")
        env["LD_LIBRARY_PATH"] = blas + os.pathsep + env.get("LD_LIBRARY_PATH", "")
        env["LD_PRELOAD"] = SHIM
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "shim_probe.py"), *args], env=env,
                         capture_output=True, text=True, timeout=600)
    line = [l for l in out.stdout.splitlines() if l.startswith("SHIM_PROBE ")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    return json.loads(line[0][len("SHIM_PROBE "):]), out.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="no CUDA device")
def test_reference_buckling_flow_runs_on_gpu_through_the_shim(ref):
    if not os.path.exists(SHIM):
        pytest.skip("shim not built (needs the reference headers)")
    base, _ = _run(False)
    gpu, log = _run(True)
    assert "[a2ds shim]" in log and "device assembly" in log   # the GPU path really ran
    e0, e1 = np.array(base["eig"]), np.array(gpu["eig"])
    assert np.all(np.array(base["err"]) < 1e-6)
    assert np.all(np.abs(e1 - e0) <= 1e-8 * np.abs(e0)), (e0, e1)
    assert abs(gpu["res_norm"] - base["res_norm"]) <= 1e-12 * base["res_norm"]
    assert abs(gpu["a_max"] - base["a_max"]) <= 1e-10 * base["a_max"]
    assert abs(gpu["a_sum"] - base["a_sum"]) <= 1e-9 * base["a_max"]
    # natural frequencies: the reference's TACSFrequencyAnalysis with K and M from the device
    f0, f1 = np.array(base["feig"])[:5], np.array(gpu["feig"])[:5]   # those nearest the shift
    assert np.all(np.array(base["ferr"])[:5] < 1e-6 * np.abs(f0))
    assert np.all(np.abs(f1 - f0) <= 1e-8 * np.abs(f0)), (f0, f1)
    # TACS_MASS_MATRIX and assembleJacobian(alpha, beta, gamma) with qddot set
    assert abs(gpu["m_max"] - base["m_max"]) <= 1e-12 * base["m_max"]
    assert abs(gpu["m_chk"] - base["m_chk"]) <= 1e-10 * base["m_max"]
    assert abs(gpu["rd_max"] - base["rd_max"]) <= 1e-12 * base["rd_max"]
    assert abs(gpu["rd_chk"] - base["rd_chk"]) <= 1e-10 * base["rd_max"]
    assert abs(gpu["ad_max"] - base["ad_max"]) <= 1e-10 * base["ad_max"]
    assert abs(gpu["ad_chk"] - base["ad_chk"]) <= 1e-9 * base["ad_max"]
    # TACS_MAT_TRANSPOSE: symmetric element matrices, so the reference's transposed assembly is
    # its normal one to rounding, and the shim takes the same device path for both
    assert base["at_vs_ad"] < 1e-13 and gpu["at_vs_ad"] < 1e-13
    assert abs(gpu["at_max"] - base["at_max"]) <= 1e-10 * base["at_max"]
    assert abs(gpu["at_chk"] - base["at_chk"]) <= 1e-9 * base["at_max"]


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="no CUDA device")
def test_reference_buckling_flow_with_quad9_shells_through_the_shim(ref):
    """TACSLinearBuckling::solve on a cylinder of TACSQuad9Shell elements: K and G assembled by
    k_assemble9 inside the unmodified flow, eigenvalues against the reference's own assembly"""
    if not os.path.exists(SHIM):
        pytest.skip("shim not built (needs the reference headers)")
    base, _ = _run(False, "quad9")
    gpu, log = _run(True, "quad9")
    assert "[a2ds shim]" in log and "device assembly" in log
    e0, e1 = np.array(base["eig"])[:4], np.array(gpu["eig"])[:4]   # the converged ones
    assert np.all(np.array(base["err"])[:4] < 1e-6)
    assert np.all(np.abs(e1 - e0) <= 1e-8 * np.abs(e0)), (e0, e1)
    assert abs(gpu["res_norm"] - base["res_norm"]) <= 1e-12 * base["res_norm"]
    assert abs(gpu["a_max"] - base["a_max"]) <= 1e-10 * base["a_max"]
    assert abs(gpu["a_chk"] - base["a_chk"]) <= 1e-9 * base["a_max"]
    # K and G in the four TACSSchurMat blocks at a fixed state, the load path the flow solved
    # for, and G at that path (weighted checksums over every entry)
    for tag in ("k", "g", "gpath"):
        assert abs(gpu[tag + "_max"] - base[tag + "_max"]) <= 1e-10 * base[tag + "_max"]
        assert abs(gpu[tag + "_chk"] - base[tag + "_chk"]) <= 1e-9 * base[tag + "_max"]
    assert abs(gpu["path_chk"] - base["path_chk"]) <= 1e-10 * base["path_max"]


@pytest.mark.gpu
@pytest.mark.skipif(not has_gpu(), reason="no CUDA device")
def test_reference_flow_with_dependent_nodes_through_the_shim(ref):
    """a reference assembler WITH dependent nodes (TACSCreator::setDependentNodes): the shim hands
    them to a2ds_set_dependent_nodes; K and G in the TACSSchurMat blocks, the Jacobian in a
    TACSParallelMat, the residual and the buckling eigenvalues against the reference's own path"""
    if not os.path.exists(SHIM):
        pytest.skip("shim not built (needs the reference headers)")
    base, _ = _run(False, "dep")
    gpu, log = _run(True, "dep")
    assert "[a2ds shim]" in log and "device assembly" in log
    e0, e1 = np.array(base["eig"])[:4], np.array(gpu["eig"])[:4]
    assert np.all(np.array(base["err"])[:4] < 1e-6)
    assert np.all(np.abs(e1 - e0) <= 1e-8 * np.abs(e0)), (e0, e1)
    assert abs(gpu["res_norm"] - base["res_norm"]) <= 1e-12 * base["res_norm"]
    assert abs(gpu["res_chk"] - base["res_chk"]) <= 1e-11 * base["res_norm"]
    for tag in ("a", "k", "g"):
        assert abs(gpu[tag + "_max"] - base[tag + "_max"]) <= 1e-10 * base[tag + "_max"]
        assert abs(gpu[tag + "_chk"] - base[tag + "_chk"]) <= 1e-9 * base[tag + "_max"]
