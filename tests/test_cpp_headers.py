"""The C++ mirror headers (include/sgslam/*.h) are meant to be compiled INSIDE the reference tree with -DSGS_WITH_OPENCV, i.e. against the real OpenCV
headers and the reference's language level (C++11).  Real OpenCV is not installed here, so they are compiled (syntax only) against a stub that has
OpenCV's real API shape where it matters -- cv::InputArray / cv::OutputArray as proxy classes without data / cols / step / ptr members
(tests/cpp/opencv_stub) -- and, separately, against the repo's own cv_compat.h.  No device needed."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = ['ORBextractor.h', 'ORBmatcher.h', 'FrameDynamic.h', 'FrameGeometry.h', 'Detector2D.h', 'Optimizer.h']


@pytest.mark.parametrize('mode', ['opencv_api_cxx11', 'compat_cxx11', 'compat_cxx17'])
def test_mirror_headers_compile(tmp_path, mode):
    src = tmp_path / 'tu.cpp'
    present = [h for h in HEADERS if os.path.exists(os.path.join(ROOT, 'include', 'sgslam', h))]
    assert 'ORBextractor.h' in present and 'ORBmatcher.h' in present
    body = ''.join('#include "sgslam/%s"\n' % h for h in present)
    body += ('void use() { ORB_SLAM2::ORBextractor ex(1000, 1.2f, 8, 20, 7); cv::Mat im(480, 640, CV_8UC1), d; std::vector<cv::KeyPoint> k; ex(im, cv::Mat(), k, d); }\n')
    src.write_text(body)
    cmd = ['g++', '-fsyntax-only', '-Wall', '-Werror=return-type', '-I', os.path.join(ROOT, 'include')]
    if mode == 'opencv_api_cxx11':
        cmd += ['-std=c++11', '-DSGS_WITH_OPENCV', '-I', os.path.join(ROOT, 'tests', 'cpp', 'opencv_stub')]
    else:
        cmd += ['-std=c++11' if mode.endswith('11') else '-std=c++17']
    r = subprocess.run(cmd + [str(src)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
