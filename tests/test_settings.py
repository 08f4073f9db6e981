"""sgs_settings_load: the hot-path keys of the reference's settings files (Examples/TUM*.yaml; read at src/Tracking.cc:53-147 and src/System.cc:160-162)."""
import ctypes as C

import numpy as np
import pytest

from pysgs import binding as B

YAML = """%YAML:1.0

#--------------------------------------------------------------------------------------------
# Camera Parameters. Adjust them!
#--------------------------------------------------------------------------------------------
Camera.fx: 517.306408
Camera.fy: 516.469215
Camera.cx: 318.643040
Camera.cy: 255.313989

Camera.k1: 0.262383
Camera.k2: -0.953104
Camera.p1: -0.005358
Camera.p2: 0.002628
Camera.k3: 1.163314

Camera.width: 640
Camera.height: 480
Camera.fps: 30.0     # frames per second
Camera.bf: 40.0
Camera.RGB: 1
ThDepth: 40.0
DepthMapFactor: 5000.0
ORBextractor.nFeatures: 1000
ORBextractor.scaleFactor: 1.2
ORBextractor.nLevels: 8
ORBextractor.iniThFAST: 20
ORBextractor.minThFAST: 7
Viewer.KeyFrameSize: 0.05
Detector2D.detection_confidence_threshold: 0.90
Detector2D.dynamic_detection_confidence_threshold: 0.01
"""


def load(path):
    s = B.Settings()
    B.check(B.lib().sgs_settings_load(str(path).encode(), C.byref(s)))
    return s


def test_keys_of_a_tum_style_file(tmp_path):
    p = tmp_path / 'TUM1.yaml'
    p.write_text(YAML)
    s = load(p)
    f32 = np.float32
    assert (s.fx, s.fy, s.cx, s.cy) == (f32(517.306408), f32(516.469215), f32(318.643040), f32(255.313989))
    assert (s.k1, s.k2, s.p1, s.p2, s.k3) == (f32(0.262383), f32(-0.953104), f32(-0.005358), f32(0.002628), f32(1.163314))
    assert (s.width, s.height, s.rgb, s.fps, s.bf) == (640, 480, 1, 30.0, 40.0)
    assert s.th_depth == f32(f32(40.0) * f32(40.0) / f32(517.306408)) and s.depth_map_factor == f32(1.0) / f32(5000.0)
    assert (s.orb.nfeatures, s.orb.nlevels, s.orb.ini_th_fast, s.orb.min_th_fast) == (1000, 8, 20, 7) and s.orb.scale_factor == f32(1.2)
    assert s.detection_confidence_threshold == f32(0.9) and s.dynamic_detection_confidence_threshold == f32(0.01)


def test_missing_keys_read_as_zero_and_depth_factor_guard(tmp_path):
    p = tmp_path / 'a.yaml'
    p.write_text('%YAML:1.0\nCamera.fx: 500\nCamera.fy: 500\nCamera.bf: 40\nDepthMapFactor: 0.0\n')
    s = load(p)
    assert s.k3 == 0 and s.orb.nfeatures == 0 and s.detection_confidence_threshold == 0 and s.th_depth == 0
    assert s.depth_map_factor == 1.0                       # |DepthMapFactor| < 1e-5 -> 1 (src/Tracking.cc:143-144)


def test_bad_files(tmp_path):
    s = B.Settings()
    assert B.lib().sgs_settings_load(str(tmp_path / 'missing.yaml').encode(), C.byref(s)) == B.SGS_ERR_INVALID
    (tmp_path / 'b.yaml').write_text('Camera.fx: 500\n')
    assert B.lib().sgs_settings_load(str(tmp_path / 'b.yaml').encode(), C.byref(s)) == B.SGS_ERR_INVALID       # no %YAML header
    (tmp_path / 'c.yaml').write_text('%YAML:1.0\nCamera.fy: 500\n')
    assert B.lib().sgs_settings_load(str(tmp_path / 'c.yaml').encode(), C.byref(s)) == B.SGS_ERR_INVALID       # no Camera.fx
