"""The reference's XCTest suite re-stated in C (tests/c/reference_tests.c) and linked against the drop-in library with gcc:
the boundary is usable from plain C exactly as the reference's header was."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "lbaudiodetective_b200")


@pytest.fixture(scope="module")
def binary(tmp_path_factory, lb):
    exe = str(tmp_path_factory.mktemp("c") / "reference_tests")
    subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "reference_tests.c"),
                    "-L", LIBDIR, "-lLBAudioDetectiveCUDA", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe], check=True)
    return exe


def test_c_program_links_and_reports_missing_device(binary, lb):
    """CPU container: the program links against the C-ABI and, with no GPU, exits with the 'skipped' status instead of computing."""
    if lb.device_available():
        pytest.skip("a CUDA device is present; see the gpu test")
    r = subprocess.run([binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_reference_suite_in_c(binary):
    r = subprocess.run([binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    print(r.stdout); print(r.stderr)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference tests passed" in r.stdout
