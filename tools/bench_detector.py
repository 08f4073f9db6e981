"""Times the detector (sgs_detector_detect_device) on synthetic RGB frames resident in HBM: frames/s per batch size, CUDA events on the launching
stream.  Usage: python tools/bench_detector.py [--size WxH] [batch ...]   (model: oracle/_ref/ncnn_model when staged, else the synthetic graph of the tests)"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, 'sg-slam_b200'), os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import torch  # noqa: E402

import detector_model as DM  # noqa: E402
from pysgs import binding as B  # noqa: E402

REAL = os.path.join(ROOT, 'oracle', '_ref', 'ncnn_model', 'mobilenetv3_ssdlite_voc')


def main():
    flags = 0
    argv = sys.argv[1:]
    W, H = 640, 480
    once = False
    if argv and argv[0] == '--once':       # one warm-up call + one call (for an ncu launch list)
        once = True; argv = argv[1:]
    layers = False
    if argv and argv[0] == '--layers':     # per-kernel table (CUDA events around every launch, sgs_detector_set_profiling): kernel list of describe() with its us per call
        layers = True; argv = argv[1:]
    if argv and argv[0] == '--size':
        W, H = [int(x) for x in argv[1].split('x')]; argv = argv[2:]
    batches = [int(a) for a in argv] or [1, 8, 64, 256]
    if os.path.exists(REAL + '.param'):
        pp, bp, name = REAL + '.param', REAL + '.bin', 'mobilenetv3_ssdlite_voc'
    else:
        pp, bp = DM.write_mini_model(tempfile.mkdtemp(), 0); name = 'synthetic-mini'
    base = np.stack([DM.synthetic_rgb(H, W, s) for s in range(8)])
    for F in batches:
        det = B.Detector(pp, bp, max_frames=F, flags=flags)
        d = torch.from_numpy(base[np.arange(F) % 8]).cuda()
        nd = torch.zeros(F, dtype=torch.int32, device='cuda'); boxes = torch.zeros((F, 4, 4), device='cuda'); have = torch.zeros(F, dtype=torch.uint8, device='cuda')
        run = lambda: det.detect_device(d.data_ptr(), H * W * 3, W * 3, W, H, F, d_dyn_rm=boxes.data_ptr(), d_ndyn_rm=nd.data_ptr(), d_have_dyn_rm=have.data_ptr(), max_boxes=4)
        for _ in range(1 if once else 3):
            run()
        torch.cuda.synchronize()
        reps = 1 if once else max(3, min(50, 2000 // F))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print('%s %dx%d batch %4d: %8.3f ms/batch  %9.1f frames/s  %6.2f TFLOP/s (1.115 GFLOP/frame)  kernels/batch %d' % (name, W, H, F, ms, F / ms * 1e3, F * 1.115 / ms, det.num_kernels), flush=True)
        if layers:
            det.set_profiling(1)
            for _ in range(6):
                run()
            torch.cuda.synchronize()
            ms, nc = det.kernel_times()
            ops = [l for l in det.describe().split('\n')[1:] if l]
            rows = [('preprocess', ms[0])] + [(o, ms[1 + j]) for j, o in enumerate(ops)] + [('detout_class', ms[-2]), ('detout_merge', ms[-1])]
            print('per-kernel us per call of %d frames (%d profiled calls), sum %.1f us' % (F, nc, sum(ms) / nc * 1e3))
            for o, t in rows:
                f = o.split(' | ')[0].split()
                short = o if len(f) < 3 else ' '.join([f[0], f[1]] + [x for x in f if '->' in x] + (f[f.index('dw') + 2:f.index('dw') + 5] if 'dw' in f else []) + (f[f.index('tile'):] if 'tile' in f else []))
                print('%9.1f  %s' % (t / nc * 1e3, short))
            det.set_profiling(0)
        det.close()


if __name__ == '__main__':
    main()
