"""The C++ host layer (include/sindyn_classes.hpp: ORB_SLAM2::DynaDetect / ORB_SLAM2::ORBextractor with the reference's
method names) and the reference-shaped driver examples/rgbd_tum_noros.cpp (rgbd_tum_noros.cc:37-215)."""
import os
import subprocess

import cv2
import numpy as np
import pytest

from sindslam_b200 import build, synth

YAML = """%YAML:1.0
Camera.fx: 535.4
Camera.fy: 539.2
Camera.cx: 320.1
Camera.cy: 247.6
Camera.RGB: 1
DepthMapFactor: 5000.0
ORBextractor.nFeatures: 1500
ORBextractor.scaleFactor: 1.2
ORBextractor.nLevels: 8
ORBextractor.iniThFAST: 15
ORBextractor.minThFAST: 5
"""


def _write(tmp_path, frames):
    root = str(tmp_path / "seq")
    synth.write_tum_sequence(root, frames, synth.TUM3, raw=True)
    y = str(tmp_path / "TUM3.yaml")
    open(y, "w").write(YAML)
    return root, y


def test_driver_builds_and_fails_loudly_without_gpu(tmp_path, lib_built):
    exe = build.build_examples()
    assert os.path.exists(exe)
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    _, frames = synth.make_sequence(2, synth.TUM3, seq=0, kind="box", start=8)
    root, y = _write(tmp_path, frames)
    r = subprocess.run([exe, "voc", y, root, os.path.join(root, "associations.txt")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2 and "sindyn error 3" in r.stderr      # SINDYN_ERR_NO_DEVICE: there is no CPU fallback


@pytest.mark.gpu
def test_driver_matches_c_abi(tmp_path, seq_c1):
    from sindslam_b200.capi import Orb, SinDyn
    exe = build.build_examples()
    _, frames = seq_c1
    root, y = _write(tmp_path, frames)
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([exe, "voc", y, root, os.path.join(root, "associations.txt"), str(out)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-800:], r.stderr[-400:])
    assert r.returncode == 0
    assert "mean dynamic detecting time" in r.stdout
    cam = synth.TUM3
    sd = SinDyn(cam.width, cam.height, cam.fx, cam.fy, cam.cx, cam.cy, cam.depth_factor)
    sd.set_prev_frames(frames[0].bgr, frames[0].bgr)
    orb = Orb(1500, 1.2, 8, 15, 5, cam.width, cam.height)
    counts = [int(l.split("keypoints")[1]) for l in r.stdout.splitlines() if "keypoints" in l]
    for k in range(1, len(frames)):
        mask, label = sd.detect(frames[k].bgr, frames[k].depth, k)
        mask = sd.morph_ellipse(mask, 15, 0)
        m = cv2.imread(str(out / ("%06d_mask.pgm" % k)), -1)
        l = cv2.imread(str(out / ("%06d_label.pgm" % k)), -1)
        assert np.array_equal(m, mask) and np.array_equal(l, label)
        gray = cv2.cvtColor(frames[k].bgr, cv2.COLOR_RGB2GRAY)     # Camera.RGB: 1 on BGR data, like the reference
        kps, _ = orb.extract(gray, mask)
        assert len(kps) == counts[k]
    sd.close()
    orb.close()


def _build_opencv_overload_check(tmp_path, lib_built):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "opencv_overloads")
    libdir = os.path.dirname(lib_built)
    cmd = ["/usr/bin/g++", "-std=c++17", "-Wall", "-Werror", "-O1", "-I" + os.path.join(root, "tests", "stubs"), "-I" + os.path.join(root, "include"),
           os.path.join(root, "tests", "stubs", "opencv_overloads.cpp"), "-o", exe, "-L" + libdir, "-lsindyn_cuda", "-Wl,-rpath," + libdir]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_reference_signature_overloads_type_check(tmp_path, lib_built):
    """include/sindyn_classes.hpp with -DSINDYN_WITH_OPENCV: the exact reference signatures -- DynaDetect(const cv::InputArray&,
    const cv::InputArray&, float x5), DetectDynaArea(const cv::InputArray&, const cv::InputArray&, cv::OutputArray&,
    cv::OutputArray&, int) (DynaDetect.h:98-131), ORBextractor::operator()(cv::InputArray, cv::InputArray,
    std::vector<cv::KeyPoint>&, cv::OutputArray) (ORBextractor.h:62-64) -- compile, warning-free, against a stand-in
    <opencv2/core.hpp> (tests/stubs; OpenCV C++ headers are not installed here) in a caller written like the reference driver.
    Without a device the program must fail loudly (exit 3), never produce outputs."""
    exe = _build_opencv_overload_check(tmp_path, lib_built)
    import torch
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "sindyn::Error" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_signature_overloads_run(tmp_path, lib_built):
    exe = _build_opencv_overload_check(tmp_path, lib_built)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("ok: mask 640x480 label 640x480"), r.stdout + r.stderr
