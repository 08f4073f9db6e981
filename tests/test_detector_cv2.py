"""The two SSD layers of the detector restatement whose semantics are not plain tensor algebra -- PriorBox and DetectionOutput -- checked against an INDEPENDENT
implementation: OpenCV's dnn module (cv2.dnn, the real library) implements the same Caffe-SSD layers that ncnn ported (ncnn's priorbox.cpp / detectionoutput.cpp and
OpenCV's prior_box_layer.cpp / detection_output_layer.cpp descend from the same Caffe SSD sources).  ncnn itself is not installable here, so this
is the closest executable pin of oracle/detector_oracle.py's prior_boxes / detection_output: decode with the prior variances (CENTER_SIZE), per-class confidence
threshold + top-k + greedy NMS, global keep-top-k; prior boxes in the order min, sqrt(min*max), then every aspect ratio with its flip, normalised corners, variances
in the second row.  What it cannot pin (stated in DESIGN.md): ncnn's two mmdetection switches of PriorBox (keys 14 / 15: stride = ceil(image / feature), first centre =
offset * (stride - 1)) -- they enter here as the step / offset handed to OpenCV -- and the order of equal scores.  No device needed."""
import os

import numpy as np
import pytest

import detector_oracle as DO
import ncnn_model as NM

cv2 = pytest.importorskip('cv2')
REAL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref', 'ncnn_model', 'mobilenetv3_ssdlite_voc')


class _Layer:                                              # the parameter view detector_oracle reads (ncnn keys)
    def __init__(self, kw): self.kw = dict(kw)
    def p(self, k, d=None): return self.kw.get(k, d)


def _net(proto):
    return cv2.dnn.readNetFromCaffe(np.frombuffer(proto.encode(), np.uint8))


def cv_detection_output(loc, conf, prior, ncls, nms, topk, keep, thr):
    P = loc.size // 4
    net = _net('''name: "t"
input: "loc" input_shape { dim: 1 dim: %d }
input: "conf" input_shape { dim: 1 dim: %d }
input: "prior" input_shape { dim: 1 dim: 2 dim: %d }
layer { name: "detection_out" type: "DetectionOutput" bottom: "loc" bottom: "conf" bottom: "prior" top: "detection_out"
  detection_output_param { num_classes: %d share_location: true background_label_id: 0 nms_param { nms_threshold: %r top_k: %d } code_type: CENTER_SIZE keep_top_k: %d confidence_threshold: %r } }
''' % (P * 4, P * ncls, P * 4, ncls, nms, topk, keep, thr))
    net.setInput(np.ascontiguousarray(loc, np.float32).reshape(1, -1), 'loc')
    net.setInput(np.ascontiguousarray(conf, np.float32).reshape(1, -1), 'conf')
    net.setInput(np.ascontiguousarray(prior, np.float32).reshape(1, 2, -1), 'prior')
    out = net.forward().reshape(-1, 7)
    return out[out[:, 1] > 0][:, 1:]                        # [label, score, xmin, ymin, xmax, ymax]; an empty result is one row of zeros


def _scene(rng, P, ncls, clustered):
    if clustered:                                           # few objects, many priors on each: NMS has real work
        k = rng.randint(0, 6, P)
        centre = rng.uniform(0.2, 0.8, (6, 2))[k] + rng.normal(0, 0.02, (P, 2)); size = rng.uniform(0.15, 0.3, (6, 2))[k] * rng.uniform(0.8, 1.25, (P, 2))
    else:
        centre = rng.uniform(0.1, 0.9, (P, 2)); size = rng.uniform(0.05, 0.4, (P, 2))
    pb = np.concatenate([centre - size / 2, centre + size / 2], 1).astype(np.float32)
    prior = np.stack([pb.reshape(-1), np.tile(np.array([0.1, 0.1, 0.2, 0.2], np.float32), P)])
    loc = rng.normal(0, 1.0, (P, 4)).astype(np.float32)
    logits = rng.normal(0, 2.5, (P, ncls)); logits[:, 0] += 1.0
    e = np.exp(logits - logits.max(1, keepdims=True)); conf = (e / e.sum(1, keepdims=True)).astype(np.float32)
    return loc, conf, prior


def _canon(rows):
    return rows[np.lexsort((rows[:, 2], rows[:, 0], -rows[:, 1]))]


def test_detection_output_equals_opencv():
    rng = np.random.RandomState(0)
    cases = [(21, 0.45, 300, 100, 0.01)]                    # the reference model's layer (param line 410)
    if os.path.exists(REAL + '.param'):
        L = [l for l in NM.parse_param(REAL + '.param') if l.type == 'DetectionOutput'][0]
        assert (L.p(0), round(L.p(1), 4), L.p(2), L.p(3), round(L.p(4), 4)) == (21, 0.45, 300, 100, 0.01)
    cases += [(21, 0.3, 50, 20, 0.05), (5, 0.6, 400, 200, 0.2), (2, 0.45, 300, 100, 0.5)]
    total = 0
    for ncls, nms, topk, keep, thr in cases:
        for trial in range(6):
            P = int(rng.choice([40, 500, 3000]))
            loc, conf, prior = _scene(rng, P, ncls, clustered=trial % 2 == 0)
            L = _Layer({0: ncls, 1: nms, 2: topk, 3: keep, 4: thr})
            mine = DO.detection_output(L, loc.reshape(-1), conf, prior)
            ref = cv_detection_output(loc, conf, prior, ncls, nms, topk, keep, thr)
            assert mine.shape == ref.shape, (ncls, nms, topk, keep, thr, trial, mine.shape, ref.shape)
            if len(mine):
                a, b = _canon(mine), _canon(ref)
                assert np.array_equal(a[:, 0], b[:, 0]) and np.abs(a[:, 1:] - b[:, 1:]).max() < 1e-6
            assert np.all(np.diff(mine[:, 1]) <= 0)         # ncnn's output order: descending score
            total += len(mine)
    assert total > 500
    # nothing above the threshold: no rows
    loc, conf, prior = _scene(rng, 100, 21, False)
    conf[:] = 0.001; conf[:, 0] = 0.98
    assert DO.detection_output(_Layer({0: 21, 1: 0.45, 2: 300, 3: 100, 4: 0.01}), loc.reshape(-1), conf, prior).shape == (0, 6)


def cv_prior_box(fw, fh, iw, ih, mins, maxs, ars, flip, clip, var, step, offset):
    proto = '''name: "t"
input: "fm" input_shape { dim: 1 dim: 1 dim: %d dim: %d }
input: "data" input_shape { dim: 1 dim: 3 dim: %d dim: %d }
layer { name: "pb" type: "PriorBox" bottom: "fm" bottom: "data" top: "pb"
  prior_box_param { %s %s %s flip: %s clip: %s %s step: %r offset: %r } }
''' % (fh, fw, ih, iw, ' '.join('min_size: %r' % m for m in mins), ' '.join('max_size: %r' % m for m in maxs), ' '.join('aspect_ratio: %r' % a for a in ars),
       'true' if flip else 'false', 'true' if clip else 'false', ' '.join('variance: %r' % v for v in var), step, offset)
    net = _net(proto)
    net.setInput(np.zeros((1, 1, fh, fw), np.float32), 'fm'); net.setInput(np.zeros((1, 3, ih, iw), np.float32), 'data')
    return net.forward().reshape(2, -1)


def test_prior_boxes_equal_opencv():
    var = [0.1, 0.1, 0.2, 0.2]
    # Caffe-style layers (ncnn keys 14 / 15 off): centre = (j + offset) * step
    for fw, fh, iw, ih, mins, maxs, ars, flip, clip, step in [(19, 19, 300, 300, [60.0], [105.0], [2.0], True, False, 16.0), (10, 10, 300, 300, [105.0], [150.0], [2.0, 3.0], True, False, 32.0),
                                                              (5, 3, 320, 200, [40.0], [], [2.0], False, True, 64.0), (3, 3, 300, 300, [195.0], [240.0], [2.0, 3.0], True, True, 100.0)]:
        L = _Layer({0: mins, 1: maxs, 2: ars, 3: var[0], 4: var[1], 5: var[2], 6: var[3], 7: int(flip), 8: int(clip), 9: iw, 10: ih, 11: step, 12: step, 13: 0.5})
        mine = DO.prior_boxes(L, fw, fh, iw, ih)
        ref = cv_prior_box(fw, fh, iw, ih, mins, maxs, ars, flip, clip, var, step, 0.5)
        assert mine.shape == ref.shape and np.abs(mine - ref).max() < 1e-6, (fw, fh, np.abs(mine - ref).max())
    # the reference model's six layers (keys 14 = 15 = 1): stride and first centre follow the mmdetection assumptions, everything else -- box order, sizes,
    # normalisation, variances -- must equal OpenCV's layer fed with that stride and the offset that yields the same centres
    if os.path.exists(REAL + '.param'):
        layers = NM.parse_param(REAL + '.param')
        fms = [19, 10, 5, 3, 2, 1]
        for L, f in zip([l for l in layers if l.type == 'PriorBox'], fms):
            assert L.p(14, 0) == 1 and L.p(15, 0) == 1
            mine = DO.prior_boxes(L, f, f, 300, 300)
            stride = float(np.ceil(300 / f)); off = float(L.p(13, 0.0)) * (stride - 1) / stride
            ref = cv_prior_box(f, f, 300, 300, list(L.p(0, [])), list(L.p(1, [])), list(L.p(2, [])), bool(L.p(7, 1)), bool(L.p(8, 0)), var, stride, off)
            assert mine.shape == ref.shape and np.abs(mine - ref).max() < 2e-6, (f, np.abs(mine - ref).max())
