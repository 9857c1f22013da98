"""CPU: the HOST logic of the compiled matcher replacements (orb_slam3_ros_b200/host/ORBmatcherGPU.cc: query construction, in-order decision
loops, rescans and fallbacks for the nine ORBmatcher searches) against the REFERENCE's own function bodies (oracle/_ref), with the device
scans replaced by a CPU test double of the few C-ABI entry points the adapter calls (tests/host/fake_orbb.cpp: the grid walk of
Frame::GetFeaturesInArea + k-nearest selection, best / second best over candidate lists).  The same comparisons run on the GPU through the
real library in tests/test_gpu_matcher_host.py; this file runs them where there is no GPU."""
import ctypes as C
import subprocess
from pathlib import Path

import pytest

from oracle import ref
import test_gpu_matcher_host as gpu_cases

ROOT = Path(__file__).resolve().parents[1]
SO = ROOT / "tests" / "models" / "_build" / "libmatcher_host_cpu.so"

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref (the reference's own object code) is not built")


@pytest.fixture(scope="module")
def host_cpu():
    SO.parent.mkdir(exist_ok=True)
    pkg = ROOT / "orb_slam3_ros_b200"
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-fPIC", "-shared", f"-I{ROOT / 'tests' / 'cvstub'}",
                           f"-I{ROOT / 'tests' / 'host' / 'slam_stub'}", f"-I{ROOT / 'oracle' / 'cvshim'}", f"-I{ROOT / 'include'}", f"-I{pkg / 'host'}",
                           str(ROOT / "tests" / "host" / "matcher_host.cpp"), str(pkg / "host" / "ORBmatcherGPU.cc"),
                           str(ROOT / "tests" / "host" / "fake_orbb.cpp"), "-o", str(SO)])
    lib = C.CDLL(str(SO))
    gpu_cases.declare(lib)
    return lib


@pytest.mark.parametrize("with_stereo,th", [(False, 1.0), (True, 1.0), (False, 3.0)])
def test_search_local_points(host_cpu, with_stereo, th):
    gpu_cases.test_search_local_points_equals_reference(host_cpu, with_stereo, th)


@pytest.mark.parametrize("stereo,direction,dense", [(False, 0, False), (True, 0, False), (True, 1, False), (True, -1, True), (False, 0, True)])
def test_motion_model(host_cpu, stereo, direction, dense):
    gpu_cases.test_motion_model_search_equals_reference(host_cpu, stereo, direction, dense)


@pytest.mark.parametrize("th,orb_dist,check,seed", [(10.0, 100, True, 5), (3.0, 64, True, 5), (10.0, 100, False, 6)])
def test_relocalization(host_cpu, th, orb_dist, check, seed):
    gpu_cases.test_relocalization_search_equals_reference(host_cpu, th, orb_dist, check, seed)


@pytest.mark.parametrize("levelsup,nnratio,check", [(2, 0.7, True), (3, 0.75, True), (2, 0.9, False)])
def test_search_by_bow(host_cpu, levelsup, nnratio, check):
    gpu_cases.test_search_by_bow_equals_reference(host_cpu, levelsup, nnratio, check)


@pytest.mark.parametrize("levelsup,nnratio,check,seed", [(2, 0.8, True, 17), (3, 0.75, True, 18), (2, 0.9, False, 19)])
def test_search_by_bow_between_key_frames(host_cpu, levelsup, nnratio, check, seed):
    gpu_cases.test_search_by_bow_between_key_frames_equals_reference(host_cpu, levelsup, nnratio, check, seed)


@pytest.mark.parametrize("crowd,jitter,window,check", [(False, 0.0, 100, True), (True, 0.0, 100, True), (False, 6.0, 40, True), (True, 3.0, 100, False)])
def test_search_for_initialization(host_cpu, crowd, jitter, window, check):
    gpu_cases.test_search_for_initialization_equals_reference(host_cpu, crowd, jitter, window, check)


@pytest.mark.parametrize("seed,th,ratio", [(5, 5, 1.0), (5, 3, 1.5), (6, 8, 1.5)])
def test_sim3_projection(host_cpu, seed, th, ratio):
    gpu_cases.test_sim3_projection_search_equals_reference(host_cpu, seed, th, ratio)


@pytest.mark.parametrize("seed,th", [(5, 4.0), (6, 4.0), (7, 8.0)])
def test_sim3_fuse(host_cpu, seed, th):
    gpu_cases.test_sim3_fuse_equals_reference(host_cpu, seed, th)


@pytest.mark.parametrize("seed,th,stereo", [(5, 3.0, False), (6, 3.0, True), (7, 6.0, True), (8, 12.0, True)])
def test_fuse(host_cpu, seed, th, stereo):
    gpu_cases.test_fuse_equals_reference(host_cpu, seed, th, stereo)


@pytest.mark.parametrize("stereo,only_stereo,coarse,check,levelsup,ties", [(False, False, False, False, 2, False), (True, False, False, True, 2, False),
                                                                           (True, True, False, False, 3, False), (False, False, True, True, 2, False),
                                                                           (False, False, True, False, 2, True), (False, False, False, False, 3, True)])
def test_search_for_triangulation(host_cpu, stereo, only_stereo, coarse, check, levelsup, ties):
    gpu_cases.test_search_for_triangulation_equals_reference(host_cpu, stereo, only_stereo, coarse, check, levelsup, ties)


@pytest.mark.parametrize("seed,th", [(41, 1.0), (42, 3.0), (43, 1.0)])
def test_search_local_points_fisheye_stereo(host_cpu, seed, th):
    gpu_cases.test_search_local_points_fisheye_stereo_equals_reference(host_cpu, seed, th)


@pytest.mark.parametrize("direction,dense,check", [(0, False, True), (1, False, True), (-1, True, True), (0, True, False)])
def test_motion_model_fisheye_stereo(host_cpu, direction, dense, check):
    """CPU test double only (see fisheye_motion_case): this branch has not run on a GPU yet"""
    gpu_cases.fisheye_motion_case(host_cpu, direction, dense, check)
