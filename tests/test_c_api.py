"""The reference's XCTest suite re-stated in C (tests/c/reference_tests.c) and linked against the drop-in library with gcc:
the boundary is usable from plain C exactly as the reference's header was."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "lbaudiodetective_b200")


def compile_c(tmp_path_factory, name):
    exe = str(tmp_path_factory.mktemp("c") / name)
    subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", name + ".c"),
                    "-L", LIBDIR, "-lLBAudioDetectiveCUDA", "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe], check=True)
    return exe


@pytest.fixture(scope="module")
def binary(tmp_path_factory, lb):
    return compile_c(tmp_path_factory, "reference_tests")


@pytest.fixture(scope="module")
def sharded_binary(tmp_path_factory, lb):
    return compile_c(tmp_path_factory, "sharded_search")


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


def test_sharded_search_program_links_and_reports_missing_device(sharded_binary, lb):
    if lb.device_available():
        pytest.skip("a CUDA device is present; see the gpu test")
    r = subprocess.run([sharded_binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_sharded_search_in_c(sharded_binary, figure):
    """A plain-C caller runs an 8-shard database search through the library alone (LBAudioDetectiveDatabaseGroup*): same top-k as
    one database, on however many GPUs the box has (the shards go round-robin over the visible devices)."""
    r = subprocess.run([sharded_binary, "60000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    figure(r.stdout.strip().splitlines()[0] if r.stdout.strip() else r.stderr.strip())
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout


@pytest.fixture(scope="module")
def sharded_extract_binary(tmp_path_factory, lb):
    return compile_c(tmp_path_factory, "sharded_extract")


def test_sharded_extract_program_links_and_reports_missing_device(sharded_extract_binary, lb):
    if lb.device_available():
        pytest.skip("a CUDA device is present; see the gpu test")
    r = subprocess.run([sharded_extract_binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 77 and "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_sharded_extract_in_c(sharded_extract_binary, figure):
    """A plain-C caller fingerprints one batch on several GPUs through the library alone: one detective per device
    (LBAudioDetectiveSetDevice), one call (LBAudioDetectiveProcessPCMBatchSharded); same words as one detective."""
    # two clips per upload chunk (a test knob read when a plan is built), so that the seven clips really are divided between the detectives
    r = subprocess.run([sharded_extract_binary], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, env={**os.environ, "LBAD_CHUNK_CLIPS": "2"})
    figure(r.stdout.strip().splitlines()[0] if r.stdout.strip() else r.stderr.strip())
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all checks passed" in r.stdout
