"""TEST INFRASTRUCTURE (CPU baseline of bench.py only): the detector restatement of detector_oracle.py evaluated on a BATCH of frames with PyTorch-CPU tensors
end to end (no per-layer numpy round trips), so that the CPU arm of the bench times a reasonably efficient CPU implementation of Detector2D::detect
(src/Detector2D.cc:34-89) on all host cores instead of a layer-by-layer interpreter on one frame.  Same graph walk and layer semantics as detector_oracle.forward
(which stays the parity checker: single frame, numpy, explicit float32 steps); tests/test_detector.py checks that both produce the same detections."""
import numpy as np
import torch
import torch.nn.functional as F

import detector_oracle as DO


def detection_output_fast(L, loc, conf, prior):
    """detector_oracle.detection_output with the greedy NMS of every class vectorised (float32 IoU matrix of the <= 300 sorted candidates, one row scan per kept
    box): identical rows -- same float32 expressions, same order -- without the interpreter-bound pair loop, which would make the CPU baseline slower than any
    compiled implementation (ncnn's is C++)."""
    ncls, nms_thr, nms_topk, keep_topk, conf_thr = L.p(0), np.float32(L.p(1, 0.05)), L.p(2, 300), L.p(3, 100), np.float32(L.p(4, 0.5))
    loc = loc.reshape(-1, 4).astype(np.float32); pb = prior[0].reshape(-1, 4); var = prior[1].reshape(-1, 4)
    conf = conf.reshape(-1, ncls).astype(np.float32)
    half = np.float32(0.5)
    pw = pb[:, 2] - pb[:, 0]; ph = pb[:, 3] - pb[:, 1]
    pcx = (pb[:, 0] + pb[:, 2]) * half; pcy = (pb[:, 1] + pb[:, 3]) * half
    cx = var[:, 0] * loc[:, 0] * pw + pcx; cy = var[:, 1] * loc[:, 1] * ph + pcy
    w = np.exp(var[:, 2] * loc[:, 2]).astype(np.float32) * pw; h = np.exp(var[:, 3] * loc[:, 3]).astype(np.float32) * ph
    boxes = np.stack([cx - w * half, cy - h * half, cx + w * half, cy + h * half], 1).astype(np.float32)
    rows = []
    hit = conf[:, 1:] > conf_thr
    for c in (np.nonzero(hit.any(axis=0))[0] + 1):
        idx = np.nonzero(hit[:, c - 1])[0]
        idx = idx[np.lexsort((idx, -conf[idx, c]))][:nms_topk]
        b = boxes[idx]
        area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
        # pair (i = later candidate b, j = earlier kept box a) exactly as the oracle evaluates it
        a = b[None, :, :]; q = b[:, None, :]
        disjoint = (q[..., 0] > a[..., 2]) | (q[..., 2] < a[..., 0]) | (q[..., 1] > a[..., 3]) | (q[..., 3] < a[..., 1])
        inter = (np.minimum(a[..., 2], q[..., 2]) - np.maximum(a[..., 0], q[..., 0])) * (np.minimum(a[..., 3], q[..., 3]) - np.maximum(a[..., 1], q[..., 1]))
        inter = np.where(disjoint, np.float32(0), inter).astype(np.float32)
        union = area[None, :] + area[:, None] - inter
        with np.errstate(divide='ignore', invalid='ignore'):
            over = (inter / union) > nms_thr
        alive = np.ones(len(idx), bool)
        picked = []
        for i in range(len(idx)):
            if alive[i]:
                picked.append(i)
                alive[i + 1:] &= ~over[i + 1:, i]
        rows += [(int(c), conf[idx[i], c], int(idx[i])) for i in picked]
    rows.sort(key=lambda r: (-r[1], r[0], r[2]))
    rows = rows[:keep_topk]
    return np.array([[c, s, *boxes[i]] for c, s, i in rows], np.float32).reshape(-1, 6)


class BatchedDetector:
    def __init__(self, layers):
        self.layers = layers
        self.w = {}
        for L in layers:
            if L.type in ('Convolution', 'ConvolutionDepthWise'):
                self.w[id(L)] = (torch.from_numpy(np.ascontiguousarray(L.weight)), torch.from_numpy(np.ascontiguousarray(L.bias)) if L.bias is not None else None)
            elif L.type == 'MemoryData':
                self.w[id(L)] = torch.from_numpy(np.ascontiguousarray(L.data))

    @torch.inference_mode()
    def forward(self, xb):
        """xb: float32 tensor [B,3,H,W] (preprocessed).  Returns (loc [B,P*4], conf [B,P,classes], priors (2, P*4) numpy, the DetectionOutput layer)."""
        blobs = {}
        det = None
        for L in self.layers:
            ins = [blobs[n] for n in L.inputs]
            ty = L.type
            if ty == 'Input':
                out = [xb]
            elif ty == 'MemoryData':
                out = [self.w[id(L)]]
            elif ty == 'Split':
                out = [ins[0]] * len(L.outputs)
            elif ty in ('Convolution', 'ConvolutionDepthWise'):
                W, b = self.w[id(L)]
                out = [F.conv2d(ins[0], W, b, stride=(L.p(13, L.p(3, 1)), L.p(3, 1)), padding=(L.p(14, L.p(4, 0)), L.p(4, 0)),
                                dilation=(L.p(12, L.p(2, 1)), L.p(2, 1)), groups=L.group)]
            elif ty == 'ReLU':
                out = [torch.relu(ins[0])]
            elif ty == 'Clip':
                out = [torch.clamp(ins[0], float(L.p(0)), float(L.p(1)))]
            elif ty == 'BinaryOp':
                a, b = ins
                if isinstance(b, torch.Tensor) and b.ndim == 1 and b.numel() == 1: b = b.reshape(())
                out = [{0: torch.add, 1: torch.sub, 2: torch.mul, 3: torch.div}[L.p(0, 0)](a, b)]
            elif ty == 'Permute':
                out = [ins[0].permute(0, 2, 3, 1).contiguous()]
            elif ty == 'Flatten':
                out = [ins[0].reshape(ins[0].shape[0], -1)]
            elif ty == 'Concat':
                if isinstance(ins[0], np.ndarray):                      # the prior boxes: constants without a batch axis
                    out = [np.concatenate(ins, axis=L.p(0, 0))]
                else:
                    out = [torch.cat(ins, dim=L.p(0, 0) + 1)]
            elif ty == 'Reshape':
                w, h = L.p(0), L.p(1, -233)
                out = [ins[0].reshape(ins[0].shape[0], h, w) if h != -233 else ins[0].reshape(ins[0].shape[0], w)]
            elif ty == 'Softmax':
                out = [torch.softmax(ins[0], dim=2)]
            elif ty == 'PriorBox':
                fm, im = ins
                out = [DO.prior_boxes(L, fm.shape[3], fm.shape[2], im.shape[3], im.shape[2])]
            elif ty == 'DetectionOutput':
                det = (L, ins[0], ins[1], ins[2])
                break
            else:
                raise NotImplementedError(ty)
            for n, o in zip(L.outputs, out):
                blobs[n] = o
        return det

    @staticmethod
    def preprocess(rgb):
        """detector_oracle.preprocess with the resize done by OpenCV when cv2 is importable (bit-identical for camera-sized frames, tests/test_detector.py:
        ncnn's from_pixels_resize is documented as OpenCV's fixed-point bilinear); the numpy restatement otherwise."""
        try:
            import cv2
            small = cv2.resize(np.ascontiguousarray(rgb), (DO.TARGET, DO.TARGET), interpolation=cv2.INTER_LINEAR)
        except ImportError:
            return DO.preprocess(rgb)
        return small.astype(np.float32).transpose(2, 0, 1) - np.asarray(DO.MEAN, np.float32).reshape(3, 1, 1)

    def detect(self, rgb_frames, det_thr=0.9, dyn_thr=0.01, chunk=16):
        """Detector2D::detect for a list of u8 HxWx3 frames: [(rows, (objects, dynamic_for_mapping, dynamic_for_rm)), ...] as detector_oracle.detect.
        The network runs on chunks of `chunk` frames in channels-last layout on PyTorch's intra-op threads (torch.set_num_threads by the caller)."""
        res = []
        for c0 in range(0, len(rgb_frames), chunk):
            fr = rgb_frames[c0:c0 + chunk]
            xb = torch.from_numpy(np.stack([self.preprocess(f) for f in fr])).contiguous(memory_format=torch.channels_last)
            L, loc, conf, prior = self.forward(xb)
            loc = loc.numpy(); conf = conf.numpy()
            for i, f in enumerate(fr):
                rows = detection_output_fast(L, loc[i], conf[i], prior)
                res.append((rows, DO.postprocess(rows, f.shape[1], f.shape[0], det_thr, dyn_thr)))
        return res
