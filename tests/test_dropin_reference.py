"""Drop-in boundary, checked on the reference's own call sites: src/Tracking.cc (the tracking thread: `new ORBextractor(...)`, `ORBmatcher matcher(0.9,true)`,
SearchByProjection(cur, last, th, mono), SearchByProjection(F, local points, th), SearchByBoW(KF, F, matches), the relocalisation SearchByProjection,
SearchForInitialization), src/LocalMapping.cc (the mapping thread: Fuse(pKF, points), SearchForTriangulation), src/LoopClosing.cc (SearchByBoW(KF, KF), SearchBySim3,
SearchByProjection(KF, Scw), Fuse(KF, Scw)) and src/Frame.cc (`(*mpORBextractorLeft)(im, cv::Mat(), mvKeys, mDescriptors)`, the scale getters, ORBmatcher::DescriptorDistance /
TH_LOW / TH_HIGH of the stereo matcher) are compiled UNMODIFIED with include/sgslam/ORBextractor.h and include/sgslam/ORBmatcher.h in place of the reference's
headers (tests/cpp/dropin_reference_pre.h).  They must compile as C++11, and every sgs_* symbol the resulting objects need must be exported by libsgs_cuda.so.
Runs where the reference tree is present (the build container); no device needed -- running the mirror on a GPU is tests/test_gpu_cpp_shim.py."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/src/sg-slam'
LIBSGS = os.path.join(ROOT, 'sg-slam_b200', 'lib', 'libsgs_cuda.so')
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, 'src', 'Tracking.cc')), reason='reference tree absent')


def _compile(src, out, defs=(), path=None, std='-std=c++11'):
    o = os.path.join(ROOT, 'oracle')
    cmd = ['g++', '-O0', std, '-fPIC', '-w', '-c', '-DSGS_WITH_OPENCV'] + list(defs) + [ '-I' + os.path.join(ROOT, 'include'), '-I' + os.path.join(o, 'tracking_shim'),
           '-I' + os.path.join(o, 'g2o_shim'), '-I' + os.path.join(o, 'frame_shim'), '-I' + os.path.join(o, 'orbmatcher_shim'), '-I' + REF, '-I' + os.path.join(REF, 'include'),
           '-include', os.path.join(ROOT, 'tests', 'cpp', 'dropin_reference_pre.h'), path or os.path.join(REF, 'src', src), '-o', out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_reference_tracking_thread_compiles_against_the_mirror_headers(tmp_path):
    needed = set()
    # the mapping thread: Fuse, SearchForTriangulation; the loop-closing thread: SearchByBoW(KF, KF), SearchBySim3, SearchByProjection(KF, Scw), Fuse(KF, Scw).
    # LoopClosing.cc is compiled as gnu++11: the reference's include/LoopClosing.h:50-51 (a map whose allocator names pair<const KeyFrame*, Sim3>) trips a
    # static_assert that today's libstdc++ enables in strict -std=c++11 mode only
    per_source = {}
    for src, defs, std in (('Tracking.cc', (), '-std=c++11'), ('Frame.cc', (), '-std=c++11'), ('LocalMapping.cc', ('-DSGS_REAL_LOCALMAPPING',), '-std=c++11'),
                           ('LoopClosing.cc', ('-DSGS_REAL_LOOPCLOSING',), '-std=gnu++11')):
        obj = str(tmp_path / (src + '.o'))
        _compile(src, obj, defs, std=std)
        per_source[src] = set(re.findall(r'\bU (sgs_[a-z0-9_]+)', subprocess.run(['nm', '-u', obj], capture_output=True, text=True).stdout))
        und = subprocess.run(['nm', '-u', obj], capture_output=True, text=True).stdout
        needed |= set(re.findall(r'\bU (sgs_[a-z0-9_]+)', und))
    obj = str(tmp_path / 'calls.o')                                   # the function mirrors (PoseOptimizationGPU, UpdateTrackInView, the dyn-reject trio) on the real Frame / MapPoint
    _compile('calls', obj, path=os.path.join(ROOT, 'tests', 'cpp', 'dropin_reference_calls.cpp'))
    needed |= set(re.findall(r'\bU (sgs_[a-z0-9_]+)', subprocess.run(['nm', '-u', obj], capture_output=True, text=True).stdout))
    for s in ('sgs_pose_optimization', 'sgs_frustum', 'sgs_lk_track', 'sgs_fundamental_ransac', 'sgs_dynreject'):
        assert s in needed, (s, sorted(needed))
    # the call sites really went through the mirror: extraction, both projection matchers, the bag-of-words matcher, the relocalisation matcher
    for s in ('sgs_extractor_create', 'sgs_extract', 'sgs_match_project_lastframe', 'sgs_match_project_localmap', 'sgs_match_bow', 'sgs_match_project_keyframe', 'sgs_search_for_initialization',
              'sgs_fuse_search', 'sgs_match_bow_keyframes'):
        assert s in needed, (s, sorted(needed))
    assert {'sgs_fuse_search', 'sgs_match_bow_keyframes'} <= per_source['LocalMapping.cc'] and {'sgs_fuse_search', 'sgs_match_bow_keyframes'} <= per_source['LoopClosing.cc']
    exported = set(re.findall(r' T (sgs_[a-z0-9_]+)', subprocess.run(['nm', '-D', '--defined-only', LIBSGS], capture_output=True, text=True).stdout))
    assert needed <= exported, sorted(needed - exported)
