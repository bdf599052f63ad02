"""Reference-style user code compiles and runs against the host mirror: Phonon / TrkPhonon (phonon.h), Field with
its functor constructor and CellVolF (field.h, problem.h), Material::vel / tau(const Phonon&), the Domain and Problem classes.
The sample program is tests/compat/reference_user_code.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "montecarlocpp_b200", "host")
SRC = os.path.join(ROOT, "tests", "compat", "reference_user_code.cpp")


def _build(tmp):
    exe = os.path.join(tmp, "reference_user_code")
    cmd = ["/usr/bin/g++", "-O1", "-std=c++17", "-fopenmp", "-Wall", "-I", HOST, SRC, "-o", exe,
           "-L", os.path.join(ROOT, "montecarlocpp_b200"), "-lmcbhost", "-lmcb",
           "-Wl,-rpath," + os.path.join(ROOT, "montecarlocpp_b200")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def test_reference_style_user_code_compiles_and_runs_host_side(matfiles, tmp_path):
    exe = _build(str(tmp_path))
    disp, relax = matfiles["silicon_small"]
    r = subprocess.run([exe, disp, relax], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert "host-side ok" in r.stdout and "BulkDomain " in r.stdout and "MultiProblem " in r.stdout


@pytest.mark.gpu
def test_reference_style_user_code_solves_on_the_gpu(matfiles, tmp_path):
    exe = _build(str(tmp_path))
    disp, relax = matfiles["silicon_small"]
    r = subprocess.run([exe, disp, relax, "gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "gpu ok: 20000 phonons" in r.stdout
