#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native SG-SLAM tracking hot path (driver contract in the task statement).

One "step" = one pass of the hot path over one batch of synthetic 640x480 frames per GPU (BASELINE.json configs[1]: "TUM
fr3/walking_xyz-shaped synthetic 640x480 stream, 1xB200, extract+match+dyn-reject"):

    ORB extract  ->  LK optical flow to the previous frame  ->  dynamic-feature rejection (boxes + epipolar)  ->  SearchByProjection(cur, last)

Frames of independent streams are sharded over ranks with no data-path collective (weak scaling); one NCCL broadcast of a shared
vocabulary-sized descriptor table happens at start-up, untimed.

  value : whole-job frames/s with all inputs resident in HBM (device-timed with CUDA events on the launching stream)
  e2e   : same metric through the C-ABI front end with HOST (pinned) buffers, H2D/D2H inside the timed region
  roofline     : dominant extractor kernel's algorithmic bytes / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline : the CPU oracle (port of the reference path, incl. LK) on this box's host cores, bounded sample

Not on the GPU in this round, hence precomputed inputs of the step: the RANSAC fundamental matrix
(src/Frame.cc:469-472; here a least-squares 8-point F from the static LK tracks, computed once on the host) and the detector boxes.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, 'sg-slam_b200'), os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)

W, H, NFEAT = 640, 480, 1000
ALG_BYTES_EXTRACT = 5_902_474          # algorithmic bytes per 640x480 frame, ORB extract
ALG_BYTES_FAST_READ = 950_532          # sum of level pixels (FAST reads every level once)
ALG_BYTES_LK_PYR = 408_000 * 6 + 403_200   # one padded pyramid + derivative planes per frame (each frame is also the previous frame of the next):
                                           # per level pixel 1 B read + 1 B write + 4 B derivative, + cv::pyrDown reading L0..L2
ALG_BYTES_LK = ALG_BYTES_LK_PYR + 1000 * 4 * 2 * 529   # + N points x 4 levels x 2 images x 23^2 window bytes (~4.2 MB + pyramid)
TH = 15.0                              # Tracking.cc:919-923 (RGB-D)
LAUNCHES_PER_STEP = 18 + 12 + 1 + 3 + 1    # extract (7 resize, FAST, quadtree, 8 blur, describe) + LK (3 pyrDown, 4 Scharr, 4 border, track) + RANSAC F + depth lookup/dyn-reject/compact + match


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line: everything libraries print there (e.g. the NCCL version banner) is sent to stderr instead."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + '\n').encode())


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# synthetic workload (S2/S4 of SURVEY 8d): `nbatch` frames = streams of consecutive S2 frames, identical bytes for CPU and GPU
# ----------------------------------------------------------------------------------------------------------------------
def make_frames(nbatch, seed, unique=32):
    from pysgs import synth
    unique = min(unique, nbatch)
    base, boxes = synth.stream_s2(unique, W, H, seed=seed)
    frames = np.zeros((nbatch, H, W), np.uint8)
    bx = np.zeros((nbatch, 4), np.float32)
    for i in range(nbatch):
        rep, k = divmod(i, unique)
        if rep == 0:
            frames[i] = base[k]; bx[i] = boxes[k]
        else:   # further streams: the same camera path seen through a cyclic shift (distinct pixels, same statistics)
            dx, dy = 7 * rep, 5 * rep
            frames[i] = np.roll(np.roll(base[k], dy, 0), dx, 1)
            bx[i] = boxes[k]; bx[i, 0] = (boxes[k, 0] + dx) % W; bx[i, 1] = (boxes[k, 1] + dy) % H
    return frames, bx, unique


def prev_index(nbatch, unique):
    """Index (inside the batch) of the previous frame of the same stream; the first frame of a stream is its own predecessor."""
    f = np.arange(nbatch, dtype=np.int32)
    return np.where(f % unique != 0, f - 1, f).astype(np.int32)


def make_track_inputs(kps, desc, counts, boxes, prev_xy, cap, point_cap, pidx):
    """Per-frame inputs of the dyn-reject + match stage (host side, untimed): person boxes, u_right from a synthetic depth plane, and
    the last-frame map points = keypoints of the previous frame back-projected with that depth.  F is NOT an input any more:
    findFundamentalMat runs inside the step, on the GPU."""
    from pysgs import synth
    import scenarios as S
    B = len(counts)
    cam = synth.TUM3
    depth = synth.depth_s1(W, H)
    ur = np.full((B, cap), -1, np.float32)
    nb = np.ones(B, np.int32); have = np.ones(B, np.uint8)
    bx = np.zeros((B, 4, 4), np.float32); bx[:, 0] = boxes
    lxyz = np.zeros((B, point_cap, 3), np.float32); ldesc = np.zeros((B, point_cap, 32), np.uint8)
    lflags = np.zeros((B, point_cap), np.uint8); loct = np.zeros((B, point_cap), np.int32); lang = np.zeros((B, point_cap), np.float32)
    ln = np.zeros(B, np.int32)
    T = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (B, 1))
    for f in range(B):
        n = counts[f]
        k = kps[f, :n]
        z = depth[np.clip(k['y'].astype(np.int64), 0, H - 1), np.clip(k['x'].astype(np.int64), 0, W - 1)]
        ur[f, :n] = k['x'] - np.float32(cam['bf']) / z
        g = pidx[f]
        m = min(counts[g], point_cap)
        kk = kps[g, :m]
        zz = depth[np.clip(kk['y'].astype(np.int64), 0, H - 1), np.clip(kk['x'].astype(np.int64), 0, W - 1)]
        lxyz[f, :m] = np.stack([(kk['x'] - cam['cx']) * zz / cam['fx'], (kk['y'] - cam['cy']) * zz / cam['fy'], zz], 1)
        ldesc[f, :m] = desc[g, :m]; loct[f, :m] = kk['octave']; lang[f, :m] = kk['angle']
        lflags[f, :m] = 1 | (2 * ((np.arange(m) % 5) != 0))                              # every 5th point is a temporal point (0 observations)
        ln[f] = m
    sf = S.scale_factors()
    return dict(ur=ur, boxes=bx, nb=nb, have=have, lxyz=lxyz, ldesc=ldesc, lflags=lflags, loct=loct, lang=lang, ln=ln, T=T, sf=sf,
                pidx=np.ascontiguousarray(pidx, np.int32))


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason sampler for the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q, '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        sm = sorted(float(r[0]) for r in rows)
        load = [v for v in sm if v >= 0.6 * sm[-1]] or sm      # "under load": idle gaps do not drag the median down
        reasons = []
        for i, name in ((3, 'hw_slowdown'), (4, 'hw_thermal_slowdown'), (5, 'sw_thermal_slowdown'), (6, 'sw_power_cap')):
            if any(r[i].lower().startswith('active') for r in rows):
                reasons.append(name)
        return {'sm_mhz': load[len(load) // 2], 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons, 'samples': len(rows),
                'power_w_max': max(float(r[2]) for r in rows if r[2].replace('.', '').isdigit())}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))), 'measured'
    except Exception:
        return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


# ----------------------------------------------------------------------------------------------------------------------
def cpu_one_frame(frames, ti, f, prev_override=None, F_override=None):
    """The CPU oracle for one frame: extract -> LK -> findFundamentalMat(RANSAC) -> dyn-reject -> SearchByProjection.
    prev_override / F_override: use the GPU's LK points / F instead of the oracle's own (LK and F agree with the GPU to a tolerance only;
    given the same LK points and F, the integer stages behind them must agree exactly)."""
    import oracle as O
    from pysgs import synth
    cam = synth.TUM3
    k, d = O.extract(frames[f])
    n = len(k)
    cur = np.stack([k['x'], k['y']], 1)
    g = int(ti['pidx'][f])
    lk = O.lk_track(frames[f], frames[g], cur)
    prev = lk if prev_override is None else prev_override[:n]
    Fm = None
    if g != f:                                              # the first frame of a stream has no previous frame: nothing is rejected
        s1, s2 = O.select_static_pairs(cur, prev, ti['boxes'][g, :ti['nb'][g]], bool(ti['have'][g]))
        Fm, _, _ = O.find_fundamental_ransac(s1, s2, 1.0, 0.99)
    Fo = Fm
    if F_override is not None:
        Fm = None if np.isnan(F_override[0]) else F_override.reshape(3, 3)
    _, keep, _, restored = O.dynreject(cur, prev, None if Fm is None else Fm.reshape(9), ti['boxes'][f, :ti['nb'][f]], bool(ti['have'][f]), NFEAT)
    sel = np.arange(n) if restored else np.nonzero(keep)[0]
    fr = O.FrameArrays(k[sel], ti['ur'][f, :n][sel], d[sel], W, H, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf'], ti['sf'])
    m = int(ti['ln'][f])
    nm, mp, nc = O.search_by_projection_last(fr, ti['T'][f].reshape(4, 4), ti['T'][f].reshape(4, 4), ti['lflags'][f, :m] & 1, ti['lxyz'][f, :m],
                                             ti['ldesc'][f, :m], (ti['lflags'][f, :m] >> 1) & 1, ti['loct'][f, :m], ti['lang'][f, :m], TH)
    return dict(n=n, nsel=len(sel), nm=nm, mp=mp, k=k, d=d, sel=sel, lk=lk, F=Fo)


def cpu_frames_per_s(frames, ti, nframes, cores):
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    O.lib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        res = list(ex.map(lambda f: cpu_one_frame(frames, ti, f), range(nframes)))   # ctypes releases the GIL inside the C++ oracle
    return nframes / (time.perf_counter() - t0), res


def detector_cpu_baseline(param_path, bin_path, frames_rgb):
    """CPU baseline of Detector2D::detect on a bounded sample: the FP32 restatement (PyTorch conv stack + numpy glue, oracle/detector_oracle.py) -- NOT ncnn,
    which is not installable here.  Returns frames/s over the sample (serial; PyTorch uses the host threads inside the convolutions)."""
    import detector_oracle as DO
    import ncnn_model as NM
    layers = NM.parse_param(param_path); NM.load_weights(layers, bin_path)
    DO.detect(layers, frames_rgb[0])                      # warm-up
    t0 = time.perf_counter()
    for f in frames_rgb:
        DO.detect(layers, f)
    return len(frames_rgb) / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's own CPU path.  The reference cannot be compiled here (needs OpenCV/Eigen/ncnn/ROS,
    DESIGN.md), so this times the CPU oracle port with all host threads, on the same workload/config."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    import oracle as O
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    frames, boxes, unique = make_frames(per_step, seed=2, unique=min(32, per_step))
    pidx = prev_index(per_step, unique)
    cap = NFEAT + 64
    kps = np.zeros((per_step, cap), O.KP_DTYPE); desc = np.zeros((per_step, cap, 32), np.uint8); counts = np.zeros(per_step, np.int32)
    prev = np.zeros((per_step, cap, 2), np.float32)
    for f in range(per_step):
        k, d = O.extract(frames[f]); counts[f] = len(k); kps[f, :len(k)] = k; desc[f, :len(k)] = d
        prev[f, :len(k)] = O.lk_track(frames[f], frames[pidx[f]], np.stack([k['x'], k['y']], 1))
    ti = make_track_inputs(kps, desc, counts, boxes, prev, cap, cap, pidx)
    for _ in range(args.warmup):
        cpu_frames_per_s(frames, ti, per_step, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_frames_per_s(frames, ti, per_step, cores)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    line = {'impl': 'reference', 'metric': 'frames/sec ORB extract+match+dyn-reject 640x480', 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
            'config': {'workload': 'S2 walking_xyz-shaped synthetic 640x480 stream, ORB 1000 features, extract + LK + findFundamentalMat(RANSAC) + dyn-reject + SearchByProjection(th=15)',
                       'frames_per_step': per_step, 'note': 'CPU oracle port of the reference path (the reference itself needs OpenCV/ROS: unbuildable here)'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': '%d frames per step x %d steps' % (per_step, args.steps)},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=512, help='frames per GPU per step (512 x 307 KB = 157 MB of input > the 126 MB L2)')
    ap.add_argument('--cpu-sample', type=int, default=0, help='frames of the cpu_baseline sample (0: two per host thread, at least 48)')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    claim_stdout()
    if args.impl == 'reference':
        return run_reference(args)

    import torch
    from pysgs import binding as B
    from pysgs import synth
    import scenarios as S

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: there is no CPU fallback for the product path')
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')   # keep stdout to the single JSON line (NCCL prints its version banner to stdout)
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    NB = args.batch
    warm = max(args.warmup, 3)
    L = B.lib()
    v = C.c_void_p

    # ---- workload + one-time set-up (untimed) -------------------------------------------------------------------------
    t_setup = time.time()
    frames, boxes, unique = make_frames(NB, seed=2 + rank)
    pidx = prev_index(NB, unique)
    sf = S.scale_factors()
    cam = B.make_camera(W, H, synth.TUM3, sf)
    trk = B.Tracker(W, H, cam, NFEAT, 1.2, 8, 20, 7, max_batch=NB, point_cap=NFEAT + 64, max_boxes=4, device=local)
    cap, pcap = trk.cap, trk.point_cap
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_frames = pin((NB, H, W), torch.uint8); h_frames.numpy()[:] = frames
    h_kps = pin((NB, cap, 28), torch.uint8); h_desc = pin((NB, cap, 32), torch.uint8); h_n = pin((NB,), torch.int32)
    d_frames = h_frames.cuda()
    d_pidx = torch.from_numpy(pidx).cuda()
    st = torch.cuda.Stream()
    L.sgs_tracker_extractor.restype = C.c_void_p
    exh = v(L.sgs_tracker_extractor(trk.h))
    torch.cuda.synchronize()
    # set-up pass on the device: extract + LK; the results feed the host-side construction of u_right / last-frame points
    B.check(L.sgs_tracker_extract_device(trk.h, v(d_frames.data_ptr()), NB, C.c_size_t(W * H), W, v(st.cuda_stream)))
    B.check(L.sgs_tracker_lk_device(trk.h, v(d_frames.data_ptr()), NB, C.c_size_t(W * H), W, v(d_pidx.data_ptr()), v(st.cuda_stream)))
    B.check(L.sgs_extractor_fetch(exh, NB, v(h_kps.data_ptr()), v(h_desc.data_ptr()), cap, v(h_n.data_ptr()), v(st.cuda_stream)))
    kps0 = h_kps.numpy().reshape(NB, cap * 28).view(B.KP_DTYPE).reshape(NB, cap).copy(); desc0 = h_desc.numpy().copy(); n0 = h_n.numpy().copy()
    pp = v(); B.check(L.sgs_tracker_prev_xy_device(trk.h, C.byref(pp)))
    prev0 = B.memcpy_d2h(np.zeros((NB, cap, 2), np.float32), pp.value)
    ti = make_track_inputs(kps0, desc0, n0, boxes, prev0, cap, pcap, pidx)
    keys_h = ['ur', 'boxes', 'nb', 'have', 'lxyz', 'ldesc', 'lflags', 'loct', 'lang', 'ln', 'T', 'pidx']
    hp = {k: torch.from_numpy(np.ascontiguousarray(ti[k])).pin_memory() for k in keys_h}
    dv = {k: t.cuda(non_blocking=True) for k, t in hp.items()}
    # shared read-only database: an ORBvoc-shaped vocabulary (k = 10, L = 6: 1,111,111 nodes x 32 B = 35.6 MB of node descriptors).
    # Rank 0 owns it; with more than one rank it reaches the others through ONE ncclBroadcast at start-up (untimed, reported).
    VOC_K, VOC_L = 10, 6
    voc_nodes = (VOC_K ** (VOC_L + 1) - 1) // (VOC_K - 1)
    voc = torch.from_numpy(synth.descriptors_s5(voc_nodes, 5) if rank == 0 else np.zeros((voc_nodes, 32), np.uint8)).cuda()
    bcast_ms = None
    if dist is not None:
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); dist.broadcast(voc, 0); e1.record(); torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    voc_parent = ((np.arange(voc_nodes, dtype=np.int64) - 1) // VOC_K).astype(np.int32); voc_parent[0] = -1        # complete k-ary tree in breadth-first node order
    voc_weight = np.zeros(voc_nodes, np.float64); voc_weight[(voc_nodes - 1) // VOC_K:] = 1.0 + (np.arange(voc_nodes - (voc_nodes - 1) // VOC_K) % 7)
    voc_h = v()
    voc_host = voc.cpu().numpy()
    B.check(L.sgs_vocabulary_create(local, VOC_K, VOC_L, voc_nodes, voc_parent.ctypes.data_as(v), voc_host.ctypes.data_as(v), voc_weight.ctypes.data_as(v), C.byref(voc_h)))
    del voc_host
    h_out = dict(kps=pin((NB, cap, 28), torch.uint8), desc=pin((NB, cap, 32), torch.uint8), ur=pin((NB, cap), torch.float32), cnt=pin((NB,), torch.int32),
                 mp=pin((NB, cap), torch.int32), nm=pin((NB,), torch.int32))
    torch.cuda.synchronize()
    log('[bench] rank %d set-up %.1fs: %d frames/step, mean %.0f keypoints/frame' % (rank, time.time() - t_setup, NB, n0.mean()))

    def track_ptrs(d):   # u_right, F (NULL: the one computed on the device), boxes, nboxes, have_dyn, last_xyz, last_desc, last_flags, last_octave, last_angle, last_n, tcw_cur, tcw_last
        return [d['ur'].data_ptr(), 0] + [d[k].data_ptr() for k in ('boxes', 'nb', 'have', 'lxyz', 'ldesc', 'lflags', 'loct', 'lang', 'ln', 'T', 'T')]

    def dev_extract():
        B.check(L.sgs_tracker_extract_device(trk.h, v(d_frames.data_ptr()), NB, C.c_size_t(W * H), W, v(st.cuda_stream)))

    def dev_lk():
        B.check(L.sgs_tracker_lk_device(trk.h, v(d_frames.data_ptr()), NB, C.c_size_t(W * H), W, v(d_pidx.data_ptr()), v(st.cuda_stream)))

    d_depth = torch.from_numpy(synth.depth_s1(W, H).astype(np.float32)).cuda()      # one synthetic depth plane shared by every frame

    def dev_stereo():   # Frame::ComputeStereoFromRGBD on the device (u_right of the unfiltered keypoints)
        B.check(L.sgs_tracker_stereo_device(trk.h, NB, v(d_depth.data_ptr()), C.c_size_t(0), W, v(st.cuda_stream)))

    def dev_fm():
        B.check(L.sgs_tracker_fundamental_device(trk.h, NB, v(dv['boxes'].data_ptr()), v(dv['nb'].data_ptr()), v(dv['have'].data_ptr()),
                                                 v(d_pidx.data_ptr()), v(st.cuda_stream)))

    def dev_track():
        ptrs = track_ptrs(dv); ptrs[0] = 0          # u_right == NULL: the one computed by dev_stereo
        B.check(L.sgs_tracker_track_device(trk.h, NB, v(0), *[v(p) for p in ptrs], C.c_float(TH), 0, 1, v(st.cuda_stream)))

    def step_host():
        # call 1: frames -> keypoints (descriptors stay on the device); call 2: LK + RANSAC F + dyn-reject + match on the resident batch
        trk.extract(h_frames.data_ptr(), NB, W * H, W, h_kps.data_ptr(), 0, h_n.data_ptr())
        B.check(L.sgs_tracker_track_lk(trk.h, NB, v(hp['pidx'].data_ptr()), *[v(p) for p in track_ptrs(hp)], C.c_float(TH), 0, 1,
                                       v(h_out['kps'].data_ptr()), v(h_out['desc'].data_ptr()), v(h_out['ur'].data_ptr()), v(h_out['cnt'].data_ptr()),
                                       v(h_out['mp'].data_ptr()), v(h_out['nm'].data_ptr())))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg (value) ----------------------------------------------------------------------------------
    with torch.cuda.stream(st):
        for _ in range(warm):
            dev_extract(); dev_lk(); dev_fm(); dev_stereo(); dev_track()
    barrier()
    B.check(L.sgs_extractor_set_profiling(exh, 1))
    L.sgs_tracker_lk.restype = C.c_void_p
    lkh = v(L.sgs_tracker_lk(trk.h))
    B.check(L.sgs_lk_set_profiling(lkh, 1))
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4 * args.steps + 1)]
    barrier()
    with torch.cuda.stream(st):
        ev[0].record(st)
        for i in range(args.steps):
            dev_extract(); ev[4 * i + 1].record(st)
            dev_lk(); ev[4 * i + 2].record(st)
            dev_fm(); ev[4 * i + 3].record(st)
            dev_stereo(); dev_track(); ev[4 * i + 4].record(st)
    barrier()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[4 * args.steps]))
    extract_ms = sum(ev[4 * i].elapsed_time(ev[4 * i + 1]) for i in range(args.steps)) / args.steps
    lk_ms = sum(ev[4 * i + 1].elapsed_time(ev[4 * i + 2]) for i in range(args.steps)) / args.steps
    fm_ms = sum(ev[4 * i + 2].elapsed_time(ev[4 * i + 3]) for i in range(args.steps)) / args.steps
    track_ms = sum(ev[4 * i + 3].elapsed_time(ev[4 * i + 4]) for i in range(args.steps)) / args.steps
    clocks = sampler.stop()
    ms5 = (C.c_double * 5)(); ncalls = C.c_int()
    B.check(L.sgs_extractor_stage_times(exh, ms5, C.byref(ncalls)))
    stage_ms = [ms5[i] / max(1, ncalls.value) for i in range(5)]
    B.check(L.sgs_extractor_set_profiling(exh, 0))
    ms2 = (C.c_double * 2)()
    B.check(L.sgs_lk_stage_times(lkh, ms2, C.byref(ncalls)))
    lk_pyr_ms, lk_track_ms = ms2[0] / max(1, ncalls.value), ms2[1] / max(1, ncalls.value)
    B.check(L.sgs_lk_set_profiling(lkh, 0))
    value = world * NB * args.steps / (total_ms * 1e-3)
    prev_dev = B.memcpy_d2h(np.zeros((NB, cap, 2), np.float32), pp.value)     # LK output of the last device step
    pF, pI = v(), v()
    B.check(L.sgs_tracker_fundamental_device_ptr(trk.h, C.byref(pF), C.byref(pI)))

    # ---- Frame::ComputeBoW of the same batch (not part of the metric: reported beside it) ------------------------------------------------
    bow_word = torch.zeros((NB, cap), dtype=torch.int32, device='cuda'); bow_w = torch.zeros((NB, cap), dtype=torch.float64, device='cuda')
    bow_node = torch.zeros((NB, cap), dtype=torch.int32, device='cuda')
    dk_, dd_, dc_, _cap = v(), v(), v(), C.c_int()
    B.check(L.sgs_extractor_results_device(exh, C.byref(dk_), C.byref(dd_), C.byref(dc_), C.byref(_cap)))

    def dev_bow():
        B.check(L.sgs_bow_transform_batch_device(voc_h, dd_, dc_, cap, NB, 4, v(bow_word.data_ptr()), v(bow_w.data_ptr()), v(bow_node.data_ptr()), v(st.cuda_stream)))
    with torch.cuda.stream(st):
        dev_bow()
        eb0, eb1 = torch.cuda.Event(True), torch.cuda.Event(True)
        eb0.record(st)
        for _ in range(5):
            dev_bow()
        eb1.record(st)
    torch.cuda.synchronize()
    bow_ms = eb0.elapsed_time(eb1) / 5

    # ---- Tracking::SearchLocalPoints of the same batch (not part of the metric): Frame::isInFrustum + SearchByProjection(F, local map, th) ------
    # local map = 4 jittered copies of the last-frame points (M = 4 x ~1000 per frame), the frame keeps the matches of the step above
    MCAP = 4 * pcap
    rs_lm = np.random.RandomState(99 + rank)
    lm_xyz = np.zeros((NB, MCAP, 3), np.float32); lm_desc = np.zeros((NB, MCAP, 32), np.uint8); lm_n = (4 * ti['ln']).astype(np.int32)
    for c4 in range(4):
        lm_xyz[:, c4 * pcap:(c4 + 1) * pcap] = ti['lxyz'] * (1 + rs_lm.normal(0, 0.002, (NB, pcap, 1))).astype(np.float32)
        lm_desc[:, c4 * pcap:(c4 + 1) * pcap] = ti['ldesc']
    for f in range(NB):          # compact the 4 copies so that the first lm_n[f] rows are the live points
        m0 = int(ti['ln'][f])
        for c4 in range(1, 4):
            lm_xyz[f, c4 * m0:(c4 + 1) * m0] = lm_xyz[f, c4 * pcap:c4 * pcap + m0]; lm_desc[f, c4 * m0:(c4 + 1) * m0] = lm_desc[f, c4 * pcap:c4 * pcap + m0]
    lm_dist = np.linalg.norm(lm_xyz, axis=2).astype(np.float32)
    lm_nrm = (lm_xyz / np.maximum(lm_dist[..., None], 1e-6)).astype(np.float32)
    lm_max = (lm_dist * 1.2 ** rs_lm.uniform(0.5, 6.5, lm_dist.shape)).astype(np.float32); lm_min = (lm_max / 1.2 ** 8).astype(np.float32)
    dlm = {k2: torch.from_numpy(a2).cuda() for k2, a2 in dict(xyz=lm_xyz, desc=lm_desc, n=lm_n, nrm=lm_nrm, mx=lm_max, mn=lm_min, obs=np.ones((NB, MCAP), np.uint8)).items()}
    lm_out = dict(inview=torch.zeros((NB, MCAP), dtype=torch.uint8, device='cuda'), px=torch.zeros((NB, MCAP), device='cuda'), py=torch.zeros((NB, MCAP), device='cuda'),
                  pxr=torch.zeros((NB, MCAP), device='cuda'), lvl=torch.zeros((NB, MCAP), dtype=torch.int32, device='cuda'), vc=torch.zeros((NB, MCAP), device='cuda'))
    rk, rd, ru, rc, rmp, rnm, rnc = v(), v(), v(), v(), v(), v(), v()
    B.check(L.sgs_tracker_results_device(trk.h, C.byref(rk), C.byref(rd), C.byref(ru), C.byref(rc), C.byref(rmp), C.byref(rnm), C.byref(rnc)))
    lm_fmp = torch.zeros((NB, cap), dtype=torch.int32, device='cuda'); lm_fobs = torch.ones((NB, cap), dtype=torch.uint8, device='cuda')
    lm_nm = torch.zeros(NB, dtype=torch.int32, device='cuda'); lm_nc = torch.zeros(NB, dtype=torch.int64, device='cuda')
    lm_fmp0 = torch.from_numpy(B.memcpy_d2h(np.zeros((NB, cap), np.int32), rmp.value)).cuda()
    lm_matcher = B.Matcher(NB, cap, MCAP, device=local)
    fa = B.FrustumBatch(); fa.cam = cam
    fa.tcw, fa.mp_xyz, fa.mp_normal, fa.mp_min_dist, fa.mp_max_dist, fa.mp_n = dv['T'].data_ptr(), dlm['xyz'].data_ptr(), dlm['nrm'].data_ptr(), dlm['mn'].data_ptr(), dlm['mx'].data_ptr(), dlm['n'].data_ptr()
    fa.point_cap, fa.viewing_cos_limit = MCAP, 0.5
    fa.mp_inview, fa.proj_x, fa.proj_y, fa.proj_xr, fa.level, fa.view_cos = [lm_out[k2].data_ptr() for k2 in ('inview', 'px', 'py', 'pxr', 'lvl', 'vc')]
    la = B.LocalMapBatch(); la.cam = cam
    la.cur_kps, la.cur_desc, la.cur_uright, la.cur_n = rk.value, rd.value, ru.value, rc.value
    la.mp_inview, la.proj_x, la.proj_y, la.proj_xr, la.level, la.view_cos = fa.mp_inview, fa.proj_x, fa.proj_y, fa.proj_xr, fa.level, fa.view_cos
    la.mp_desc, la.mp_obs, la.mp_n, la.th, la.nnratio, la.id_base = dlm['desc'].data_ptr(), dlm['obs'].data_ptr(), dlm['n'].data_ptr(), 3.0, 0.8, 100000
    la.f_mp, la.f_mp_obs, la.nmatches, la.ncand = lm_fmp.data_ptr(), lm_fobs.data_ptr(), lm_nm.data_ptr(), lm_nc.data_ptr()

    def dev_localmap():
        lm_fmp.copy_(lm_fmp0)          # the frame starts from the matches of SearchByProjection(cur, last)
        B.check(L.sgs_frustum_batch_device(C.byref(fa), NB, v(st.cuda_stream)))
        lm_matcher.localmap_batch(la, NB, st.cuda_stream)
    with torch.cuda.stream(st):
        dev_localmap()
        el0, el1 = torch.cuda.Event(True), torch.cuda.Event(True)
        el0.record(st)
        for _ in range(5):
            dev_localmap()
        el1.record(st)
    torch.cuda.synchronize()
    localmap_ms = el0.elapsed_time(el1) / 5
    lm_stats = (float(lm_out['inview'].float().sum(1).mean().item()), float(lm_nm.float().mean().item()))

    # ---- Optimizer::PoseOptimization of the same batch on the matches of the step (not part of the metric) -------------------------------------
    po = B.PoseOptBatch(); po.cam = cam
    po_T = torch.zeros((NB, 16), device='cuda'); po_out = torch.zeros((NB, cap), dtype=torch.uint8, device='cuda'); po_nin = torch.zeros(NB, dtype=torch.int32, device='cuda')
    po_err = torch.zeros((NB, cap, 3), dtype=torch.float64, device='cuda'); po_lvl = torch.zeros((NB, cap), dtype=torch.uint8, device='cuda')
    po.tcw_in, po.kps, po.uright, po.n, po.cap = dv['T'].data_ptr(), rk.value, ru.value, rc.value, cap
    po.has_mp, po.mp_index, po.points_xyz, po.point_cap = 0, rmp.value, dv['lxyz'].data_ptr(), pcap
    for l in range(8):
        po.inv_level_sigma2[l] = float(1.0 / (sf[l] * sf[l]))
    po.tcw_out, po.outlier, po.ninliers, po.scratch_err, po.scratch_level = po_T.data_ptr(), po_out.data_ptr(), po_nin.data_ptr(), po_err.data_ptr(), po_lvl.data_ptr()

    def dev_poseopt():
        B.check(L.sgs_pose_optimization_batch_device(C.byref(po), NB, v(st.cuda_stream)))
    with torch.cuda.stream(st):
        dev_poseopt()
        ep0, ep1 = torch.cuda.Event(True), torch.cuda.Event(True)
        ep0.record(st)
        for _ in range(5):
            dev_poseopt()
        ep1.record(st)
    torch.cuda.synchronize()
    poseopt_ms = ep0.elapsed_time(ep1) / 5
    poseopt_inl = float(po_nin.float().mean().item()); poseopt_edges = float((lm_fmp0 >= 0).sum(1).float().mean().item())

    # ---- Detector2D::detect of the same batch (not part of the metric: the step takes the boxes as inputs, as the reference's tracking thread does) --
    # frames = the grey frames replicated to 3 channels; model = the reference's MobileNetV3-SSDLite graph when build() staged it, else the tests' synthetic graph
    det_info = None
    det_model = None
    try:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import detector_model as DM
        real = os.path.join(ROOT, 'oracle', '_ref', 'ncnn_model', 'mobilenetv3_ssdlite_voc')
        if os.path.exists(real + '.param'):
            dpp, dbp, dname, gflop = real + '.param', real + '.bin', 'mobilenetv3_ssdlite_voc (9.7 MB FP32 weights)', 1.115
        else:
            import tempfile
            dpp, dbp = DM.write_mini_model(tempfile.mkdtemp(), 0); dname, gflop = 'synthetic graph of tests/detector_model.py (reference model not staged)', None
        det_model = (dpp, dbp)
        DB = 128                                     # frames per detector call
        det = B.Detector(dpp, dbp, max_frames=DB, det_thr=0.9, dyn_thr=0.01, device=local)
        d_rgb = d_frames[:DB].reshape(DB, H, W, 1).expand(DB, H, W, 3).contiguous()
        det_boxes = torch.zeros((DB, 4, 4), device='cuda'); det_nb = torch.zeros(DB, dtype=torch.int32, device='cuda'); det_have = torch.zeros(DB, dtype=torch.uint8, device='cuda')

        def dev_detect():
            det.detect_device(d_rgb.data_ptr(), H * W * 3, W * 3, W, H, DB, d_dyn_rm=det_boxes.data_ptr(), d_ndyn_rm=det_nb.data_ptr(), d_have_dyn_rm=det_have.data_ptr(),
                              max_boxes=4, stream=st.cuda_stream)
        with torch.cuda.stream(st):
            for _ in range(2):
                dev_detect()
            ed0, ed1 = torch.cuda.Event(True), torch.cuda.Event(True)
            ed0.record(st)
            for _ in range(5):
                dev_detect()
            ed1.record(st)
        torch.cuda.synchronize()
        det_ms = ed0.elapsed_time(ed1) / 5
        pk, _ = measured_peaks()
        det_info = {'model': dname, 'frames_per_call': DB, 'ms_per_call': det_ms, 'frames_per_s': DB / det_ms * 1e3, 'kernels_per_call': det.num_kernels,
                    'note': 'Detector2D::detect (resize + 408-layer ncnn graph + DetectionOutput + box post-processing) on device-resident RGB frames, timed separately, not part of value'}
        if gflop:
            tf = DB * gflop / det_ms
            det_info['roofline'] = {'bound': 'tensor', 'achieved': tf, 'peak': pk['bf16_tflops'], 'unit': 'TFLOP/s', 'frac': tf / pk['bf16_tflops'],
                                    'note': '1.115 GFLOP per 300x300 inference (SURVEY 8d); 1x1 convolutions = error-compensated TF32 on the tensor cores (three mma.sync per product, FP32-grade accuracy), the rest FP32 FMA on the CUDA cores; peak = measured dense bf16'}
        det.close()
        del d_rgb
    except Exception as ex:      # the detector is reported, never allowed to take the headline measurement down
        log('[bench] detector stage skipped: %r' % (ex,))

    step_host()   # e2e warm-up; its outputs are also used for the parity spot-check below
    counts_after = h_out['cnt'].numpy().copy(); nmatch = h_out['nm'].numpy().copy()
    F_dev = B.memcpy_d2h(np.zeros((NB, 9), np.float64), pF.value); F_info = B.memcpy_d2h(np.zeros((NB, 4), np.int32), pI.value)

    # ---- e2e leg: host buffers through the C ABI, copies inside the timed region ----------------------------------------
    # (a) one tracker handle, synchronous calls back to back; (b) the same work split over two handles driven by two host threads,
    # so that the copies of one half overlap the kernels of the other (what a multi-camera server does).  (b) is the headline e2e.
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt1 = max_over_ranks(time.perf_counter() - t0)
        h2d = h_frames.numel() + sum(hp[k].numel() * hp[k].element_size() for k in keys_h) + hp['T'].numel() * 4
        d2h = (h_kps.numel() + h_n.numel() * 4) + sum(t.numel() * t.element_size() for t in h_out.values())
        dt, mode = dt1, 'one tracker handle, synchronous calls'
        if NB % (2 * unique) == 0 and NB >= 128:
            HB = NB // 2
            halves = []
            for hx in range(2):
                tk = B.Tracker(W, H, cam, NFEAT, 1.2, 8, 20, 7, max_batch=HB, point_cap=NFEAT + 64, max_boxes=4, device=local)
                sl = slice(hx * HB, (hx + 1) * HB)
                hpi = {k: hp[k][sl] for k in keys_h}
                hpi['pidx'] = torch.from_numpy(np.ascontiguousarray(pidx[sl] - hx * HB)).pin_memory()
                halves.append((tk, sl, hpi))

            def half_step(hx):
                tk, sl, hpi = halves[hx]
                tk.extract(h_frames[sl].data_ptr(), HB, W * H, W, h_kps[sl].data_ptr(), 0, h_n[sl].data_ptr())
                B.check(L.sgs_tracker_track_lk(tk.h, HB, v(hpi['pidx'].data_ptr()), *[v(p) for p in track_ptrs(hpi)], C.c_float(TH), 0, 1,
                                               v(h_out['kps'][sl].data_ptr()), v(h_out['desc'][sl].data_ptr()), v(h_out['ur'][sl].data_ptr()),
                                               v(h_out['cnt'][sl].data_ptr()), v(h_out['mp'][sl].data_ptr()), v(h_out['nm'][sl].data_ptr())))

            def worker(hx, nsteps, gate):
                torch.cuda.set_device(local)
                gate.wait()
                for _ in range(nsteps):
                    half_step(hx)

            def run_pair(nsteps):
                gate = threading.Barrier(3)
                th = [threading.Thread(target=worker, args=(hx, nsteps, gate)) for hx in range(2)]
                for t in th:
                    t.start()
                barrier()
                gate.wait()
                t0 = time.perf_counter()
                for t in th:
                    t.join()
                torch.cuda.synchronize()
                return time.perf_counter() - t0
            ref_cnt = h_out['cnt'].numpy().copy(); ref_nm = h_out['nm'].numpy().copy()
            run_pair(2)
            same = bool(np.array_equal(ref_cnt, h_out['cnt'].numpy()) and np.array_equal(ref_nm, h_out['nm'].numpy()))
            dt2 = max_over_ranks(run_pair(args.steps))
            if not same:
                log('[bench] WARNING: two-handle e2e results differ from the single-handle ones')
            elif dt2 < dt1:
                dt, mode = dt2, 'two tracker handles (%d frames each per call) driven by two host threads' % HB
            for tk, _, _ in halves:
                tk.close()
        e2e = {'value': world * NB * args.steps / dt, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'ms_per_step': 1e3 * dt / args.steps, 'mode': mode, 'single_handle_value': world * NB * args.steps / dt1,
               'note': 'sgs_tracker_extract (host frames -> host keypoints) + sgs_tracker_track_lk (LK + RANSAC F + dyn-reject + match on the resident batch) with pinned host buffers'}

    # ---- roofline of the dominant kernel of the step (all timed live with CUDA events on the launching stream) -----------------
    peaks, peak_kind = measured_peaks()
    names = ['pyramid(7 launches)', 'fast_warp_cells_kernel', 'quadtree_kernel', 'blur(8 launches)', 'describe_kernel']
    ncand_frame0 = 0
    for l in range(8):
        nn = C.c_int()
        L.sgs_extractor_read_candidates(exh, 0, l, None, 0, C.byref(nn))   # count only (returns SGS_ERR_CAPACITY by design)
        ncand_frame0 += nn.value
    nk = int(n0.mean())
    # algorithmic bytes per frame: extractor stages as listed there; LK tracker = N points x 4 levels x 2 images x 23^2 B
    alg = {'pyramid(7 launches)': 1_569_878, 'fast_warp_cells_kernel': ALG_BYTES_FAST_READ + 4 * ncand_frame0, 'quadtree_kernel': 8 * ncand_frame0 + 4 * nk,
           'blur(8 launches)': 1_901_064, 'describe_kernel': (749 + 544 + 60) * nk, 'lk_pyramid+deriv(11 launches)': ALG_BYTES_LK_PYR,
           'lk_track_kernel': nk * 4 * 2 * 529, 'fm_ransac_kernel': 16 * nk + 72, 'stereo+dynreject+compact+match(4 launches)': 76 * nk + 56 * nk + 44 * 8 * nk}
    all_ms = dict(zip(names, stage_ms))
    all_ms.update({'lk_pyramid+deriv(11 launches)': lk_pyr_ms, 'lk_track_kernel': lk_track_ms, 'fm_ransac_kernel': fm_ms, 'stereo+dynreject+compact+match(4 launches)': track_ms})
    dom = max(('fast_warp_cells_kernel', 'quadtree_kernel', 'describe_kernel', 'lk_track_kernel', 'fm_ransac_kernel'), key=lambda k: all_ms[k])   # single-launch kernels
    dom_bytes = alg[dom] * NB
    achieved = dom_bytes / (all_ms[dom] * 1e-3) / 1e9
    traffic = None
    try:        # dram__bytes_read + write of that kernel from the committed ncu --set full capture (profiles/), per launch
        txt = open(os.path.join(ROOT, 'profiles', 'r01d_ncu_full_summary.txt')).read().split('=== ')
        blk = [b for b in txt if b.startswith(dom)][0]
        rd = float([l for l in blk.splitlines() if 'dram__bytes_read.sum' in l][0].split()[-1]); wr = float([l for l in blk.splitlines() if 'dram__bytes_write.sum' in l][0].split()[-1])
        traffic = int((rd + wr) * 1e6)          # the summary prints Mbyte for a 512-frame launch
    except Exception:
        pass
    step_alg_bytes = (ALG_BYTES_EXTRACT + ALG_BYTES_LK + 16 * nk + 76 * nk + 56 * nk + 44 * 8 * nk) * NB
    roofline = {'bound': 'hbm', 'kernel': dom, 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                'traffic': traffic, 'peak_kind': ('measured copy bandwidth (MEASURED_PEAKS.json)' if peak_kind == 'measured' else 'fallback 6650 GB/s'),
                'algorithmic_bytes_per_launch': int(dom_bytes), 'kernel_ms': all_ms[dom],
                'stage_ms': {k: round(x, 4) for k, x in all_ms.items()}, 'extract_ms': extract_ms, 'lk_ms': lk_ms, 'fundamental_ransac_ms': fm_ms,
                'dynreject_match_ms': track_ms,
                'per_kernel_alg_gbs': {k: round(alg[k] * NB / (all_ms[k] * 1e-3) / 1e9, 1) for k in all_ms},
                'extract_alg_gbs': ALG_BYTES_EXTRACT * NB / (extract_ms * 1e-3) / 1e9,
                'extract_frac_of_hbm': ALG_BYTES_EXTRACT * NB / (extract_ms * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'step_alg_gbs': step_alg_bytes / (total_ms / args.steps * 1e-3) / 1e9,
                'step_frac_of_hbm': step_alg_bytes / (total_ms / args.steps * 1e-3) / 1e9 / peaks['hbm_gbs'],
                'note': 'every kernel of the step is instruction-issue / latency bound (DRAM throughput 0.4-18 % in profiles/): the HBM fraction is reported as asked, the binding roof is the integer ALU / issue rate (lk_track_kernel: 55 % issue-active at 16 warps/SM)'}

    # ---- CPU baseline on this box's host cores (rank 0 only, N=1 only) + parity of the sample -------------------------------
    cpu = None
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        ns = min(args.cpu_sample if args.cpu_sample > 0 else max(48, 2 * cores), NB)        # every host thread gets work: ~0.1 s of CPU per frame
        cpu_fps, res = cpu_frames_per_s(frames, ti, ns, cores)
        res = res[:48]                                                                         # parity is spot-checked on the first 48 frames of the sample
        kps_h = h_out['kps'].numpy().reshape(NB, cap * 28).view(B.KP_DTYPE).reshape(NB, cap)
        ok = True
        lk_err = []
        f_err = []
        for f, r in enumerate(res):
            n = r['n']
            ok &= int(n0[f]) == n and kps0[f, :n].tobytes() == r['k'].tobytes() and bool(np.array_equal(desc0[f, :n], r['d']))   # extraction: bit-exact
            lk_err.append(np.abs(prev_dev[f, :n] - r['lk']).max(1))                                                                 # LK: tolerance
            rr = cpu_one_frame(frames, ti, f, prev_override=prev_dev[f], F_override=F_dev[f])                                       # F given the GPU's LK; integer stages given the GPU's LK and F
            if rr['F'] is None:
                ok &= bool(np.isnan(F_dev[f, 0]))
            else:
                f_err.append(float(np.abs(rr['F'].reshape(9) - F_dev[f]).max() / max(1.0, np.abs(rr['F']).max()))); ok &= f_err[-1] <= 1e-9
            ok &= int(counts_after[f]) == rr['nsel'] and int(nmatch[f]) == rr['nm'] and kps_h[f, :rr['nsel']].tobytes() == rr['k'][rr['sel']].tobytes()
            ok &= bool(np.array_equal(h_out['desc'].numpy()[f, :rr['nsel']], rr['d'][rr['sel']])) and bool(np.array_equal(h_out['mp'].numpy()[f, :rr['nsel']], rr['mp']))
        e = np.concatenate(lk_err)
        cpu = {'value': cpu_fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': '%d frames of the same batch, %d worker threads' % (ns, cores),
               'parity_with_gpu_on_sample': bool(ok), 'lk_abs_err_px': {'median': float(np.median(e)), 'p99': float(np.quantile(e, 0.99)), 'max': float(e.max())},
               'F_rel_err_max': (max(f_err) if f_err else None)}
        if not ok:
            log('[bench] WARNING: GPU results differ from the oracle on the CPU sample')
        if det_info is not None and det_model is not None:
            try:
                sample = [np.repeat(frames[f][:, :, None], 3, axis=2) for f in range(3)]
                det_info['cpu_baseline'] = {'value': detector_cpu_baseline(det_model[0], det_model[1], sample), 'unit': 'frames/s', 'cores': os.cpu_count() or 1, 'kind': 'port',
                                            'sample': '3 frames, serial; PyTorch-CPU FP32 convolutions + numpy glue (the restatement used as the test oracle, not ncnn)'}
            except Exception as ex:
                log('[bench] detector CPU baseline skipped: %r' % (ex,))

    if rank == 0:
        line = {'metric': 'frames/sec ORB extract+match+dyn-reject 640x480', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': warm, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
                'data': 'synthetic',
                'config': {'workload': 'S2 walking_xyz-shaped synthetic 640x480 stream (BASELINE configs[1]), ORB 1000 features / 8 levels / 1.2: extract + LK(21x21, 4 levels) + findFundamentalMat(RANSAC 1.0/0.99) + dyn-reject(boxes + epipolar) + SearchByProjection(th=15)',
                           'frames_per_gpu_per_step': NB, 'l2_policy': 'inputs larger than L2: %d frames x 307200 B = %.0f MB per step (+ %.0f MB pyramid traffic)' % (NB, NB * 0.3072, NB * 0.95),
                           'sharding': 'independent streams per rank, no data-path collective; one untimed ncclBroadcast of the vocabulary node descriptors (35.6 MB) at start-up',
                           'mean_keypoints': float(n0.mean()), 'mean_after_dynreject': float(counts_after.mean()), 'mean_matches': float(nmatch.mean()),
                           'ransac_iterations_mean': float(F_info[:, 2].mean()), 'ransac_inlier_ratio_mean': float((F_info[:, 1] / np.maximum(1, F_info[:, 0])).mean()),
                           'search_local_points_ms_per_step': localmap_ms, 'search_local_points_note': 'Frame::isInFrustum + SearchByProjection(F, local map of %d points/frame, th=3): %.0f points in view, %.0f new matches per frame; timed separately, not part of value' % (int(lm_n.mean()), lm_stats[0], lm_stats[1]),
                           'pose_optimization_ms_per_step': poseopt_ms, 'pose_optimization_note': 'Optimizer::PoseOptimization on the matches of the step (%.0f of %.0f edges kept per frame); timed separately, not part of value' % (poseopt_inl, poseopt_edges),
                           'bow_transform_ms_per_step': bow_ms, 'bow_note': 'Frame::ComputeBoW (DBoW2 transform, k=10 L=6 vocabulary of %d nodes) of the same %d frames, timed separately, not part of value' % (voc_nodes, NB),
                           'detector': det_info,
                           'not_in_step': 'the detector (boxes are inputs of the step; Detector2D::detect is timed separately under config.detector)'},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': LAUNCHES_PER_STEP * args.steps, 'roofline': roofline, 'cpu_baseline': cpu}
        if bcast_ms is not None:
            line['config']['startup_broadcast_ms'] = bcast_ms
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
