"""The C++ host adapter (orb_slam3_ros_b200/host: namespace ORB_SLAM3, class ORBextractor with the reference's
interface) compiled with the reference's language level (C++14) against a stand-in for the OpenCV types, exercised with
the reference's own call sequence (Tracking.cc:631, Frame.cc:110-116, :418-425, :818-923).  CPU: it compiles and links.
GPU: its output equals what the C ABI returns through the python binding."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
EXE = ROOT / "tests" / "models" / "_build" / "adapter_check"


def _build():
    from orb_slam3_ros_b200 import build
    build.build_library()
    EXE.parent.mkdir(exist_ok=True)
    pkg = ROOT / "orb_slam3_ros_b200"
    cmd = ["g++", "-std=c++14", "-O2", f"-I{ROOT / 'tests' / 'cvstub'}", f"-I{ROOT / 'include'}", f"-I{pkg / 'host'}",
           str(ROOT / "tests" / "host" / "adapter_check.cpp"), str(pkg / "host" / "ORBextractor.cc"), f"-L{pkg}", "-lorbb200",
           f"-Wl,-rpath,{pkg}", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(EXE)]
    subprocess.check_call(cmd)


def test_adapter_compiles_and_links_as_cxx14():
    _build()
    assert EXE.exists()


def test_resolve_in_order_equals_the_sequential_reference_loop():
    """ORBmatcherGPU::ResolveInOrder (pure host code): one batched best-two scan + the in-order decision loop must give what the
    reference's scan-decide-update loop (ORBmatcher.cc:77-141) gives, on random candidate lists competing for the same key points"""
    from orb_slam3_ros_b200 import build
    build.build_library()
    exe = EXE.parent / "resolve_check"
    exe.parent.mkdir(exist_ok=True)
    pkg = ROOT / "orb_slam3_ros_b200"
    subprocess.check_call(["g++", "-std=c++14", "-O2", f"-I{ROOT / 'tests' / 'cvstub'}", f"-I{ROOT / 'include'}", f"-I{pkg / 'host'}",
                           str(ROOT / "tests" / "host" / "resolve_check.cpp"), f"-L{pkg}", "-lorbb200", f"-Wl,-rpath,{pkg}",
                           "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "resolve_check OK" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_adapter_matches_c_abi(tmp_path):
    from orb_slam3_ros_b200 import synth
    from orb_slam3_ros_b200.extractor import ORBextractor
    from orb_slam3_ros_b200.matcher import ORBmatcher
    _build()
    img = synth.frame(240, 320, 21)
    raw = tmp_path / "img.raw"
    img.tofile(raw)
    out = subprocess.run([str(EXE), str(raw), "320", "240"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    got = dict(kv.split("=") for kv in out.stdout.split())
    ge = ORBextractor(300, 1.2, 4, 20, 7)
    mono, k, d = ge(img, None, (0, 1000))
    s = 0
    M = (1 << 64) - 1
    for i in range(len(k)):
        s = (s * 1000003 + int(np.float32(k["x"][i]) * np.float32(16)) + 7 * int(np.float32(k["y"][i]) * np.float32(16)) + 13 * int(k["octave"][i])
             + 17 * int(k["response"][i]) + 19 * int(k["size"][i])) & M
        for b in d[i]:
            s = (s * 31 + int(b)) & M
    p1 = ge.image_pyramid(1)
    win = p1[10:21, 12:23].astype(np.uint64)
    psum = int((win * (np.arange(11)[:, None] * 11 + np.arange(11)[None, :] + 1).astype(np.uint64)).sum())
    assert int(got["n"]) == len(k) and int(got["mono"]) == mono and got["desc"] == f"{len(k)}x32"
    assert int(got["sum"]) == s
    assert got["pyr1"] == f"{p1.shape[1]}x{p1.shape[0]}" and int(got["psum"]) == psum
    assert int(got["d01"]) == ORBmatcher.DescriptorDistance(d[0], d[1])
    # the extensions: colour input, rectified extraction, keypoint undistortion -- against the python binding of the same C ABI
    from orb_slam3_ros_b200.rectify import Rectifier

    def ksum(k, d):
        s = 0
        for i in range(len(k)):
            s = (s * 1000003 + int(np.float32(k["x"][i]) * np.float32(16)) + 7 * int(np.float32(k["y"][i]) * np.float32(16)) + 13 * int(k["octave"][i])) & M
            for b in d[i]:
                s = (s * 31 + int(b)) & M
        return s
    bgr = np.stack([img, 255 - img, img // 2], 2)
    _, kc, dc = ge.extract_color(bgr, rgb=False)
    assert int(got["nC"]) == len(kc) and int(got["sumC"]) == ksum(kc, dc)
    yy, xx = np.mgrid[0:240, 0:320].astype(np.float32)
    r = Rectifier(xx + np.float32(2.25), yy - np.float32(1.5), img.shape)
    _, kr, dr = r.extract(ge, img)
    assert int(got["nR"]) == len(kr) and int(got["sumR"]) == ksum(kr, dr)
    un = ORBmatcher().undistort_points(np.stack([k["x"], k["y"]], 1), (458.654, 457.296, 367.215, 248.375),
                                       [-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
    s = 0
    for a, b in un.view(np.uint32):
        s = (s * 1000003 + int(a) + 7 * int(b)) & M
    assert int(got["sumU"]) == s


def test_reference_call_site_text_compiles_against_the_adapter_header(tmp_path):
    """Frame::ExtractORB (Frame.cc:418-425) and Frame::ComputeStereoMatches (Frame.cc:811-981), the reference's own text cut out at test
    time, compile against orb_slam3_ros_b200/host/ORBextractor.h (with the OpenCV stand-in of oracle/cvshim, which carries cv::norm and
    the Mat views ComputeStereoMatches needs).  Only where /root/reference exists (the development container)."""
    import sys
    ref_root = Path("/root/reference/orb_slam3")
    if not (ref_root / "src" / "Frame.cc").exists():
        pytest.skip("the reference sources are not present on this machine")
    sys.path.insert(0, str(ROOT / "oracle"))
    import cut_reference
    cut = tmp_path / "cut"
    cut.mkdir()
    lines = (ref_root / "src" / "Frame.cc").read_text(errors="replace").splitlines(keepends=True)
    (cut / "Frame_ExtractORB.inc").write_text(cut_reference.cut(lines, r"^void Frame::ExtractORB\(int flag, const cv::Mat &im", "function", "Frame::ExtractORB"))
    (cut / "Frame_ComputeStereoMatches.inc").write_text(cut_reference.cut(lines, r"^void Frame::ComputeStereoMatches\(\)", "function", "Frame::ComputeStereoMatches"))
    pkg = ROOT / "orb_slam3_ros_b200"
    r = subprocess.run(["g++", "-std=c++14", "-fsyntax-only", "-w", f"-I{ROOT / 'oracle' / 'cvshim'}", f"-I{pkg / 'host'}", f"-I{ROOT / 'include'}", f"-I{tmp_path}",
                        str(ROOT / "tests" / "host" / "callsite_check.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_latency_tool_and_matcher_harness_compile():
    """bench.py's `single_frame_cpp_adapter` leg (tests/host/adapter_latency.cpp) and the matcher harness of the GPU tests
    (tests/host/matcher_host.cpp + host/ORBmatcherGPU.cc against tests/host/slam_stub) must compile where there is no GPU"""
    from orb_slam3_ros_b200 import build
    build.build_library()
    pkg = ROOT / "orb_slam3_ros_b200"
    out = EXE.parent
    out.mkdir(exist_ok=True)
    subprocess.check_call(["g++", "-std=c++14", "-O1", f"-I{ROOT / 'tests' / 'cvstub'}", f"-I{ROOT / 'include'}", f"-I{pkg / 'host'}",
                           str(ROOT / "tests" / "host" / "adapter_latency.cpp"), str(pkg / "host" / "ORBextractor.cc"), f"-L{pkg}", "-lorbb200",
                           f"-Wl,-rpath,{pkg}", "-L/usr/local/cuda/lib64", "-lcudart", "-o", str(out / "adapter_latency_cpu_check")])
    subprocess.check_call(["g++", "-std=c++14", "-O1", "-ffp-contract=off", "-fPIC", "-shared", f"-I{ROOT / 'tests' / 'cvstub'}",
                           f"-I{ROOT / 'tests' / 'host' / 'slam_stub'}", f"-I{ROOT / 'oracle' / 'cvshim'}", f"-I{ROOT / 'include'}", f"-I{pkg / 'host'}",
                           str(ROOT / "tests" / "host" / "matcher_host.cpp"), str(pkg / "host" / "ORBmatcherGPU.cc"), f"-L{pkg}", "-lorbb200",
                           f"-Wl,-rpath,{pkg}", "-o", str(out / "libmatcher_host_cpu_check.so")])
