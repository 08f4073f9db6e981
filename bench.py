#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native SG-SLAM tracking hot path (driver contract in the task statement).

One "step" = one pass of the hot path (ORB extract -> dynamic-feature rejection -> SearchByProjection against the last
frame) over one batch of synthetic 640x480 frames per GPU (BASELINE.json configs[1]: "TUM fr3/walking_xyz-shaped synthetic
640x480 stream, 1xB200, extract+match+dyn-reject").  Frames of independent streams are sharded over ranks with no
data-path collective (weak scaling); one NCCL broadcast of the shared last-frame map database happens at start-up, untimed.

  value : whole-job frames/s with all inputs resident in HBM (device-timed with CUDA events on the launching stream)
  e2e   : same metric through the C-ABI front end with HOST (pinned) buffers, H2D/D2H inside the timed region
  roofline     : dominant kernel's algorithmic bytes / its CUDA-event time vs the measured HBM copy peak
  cpu_baseline : the CPU oracle (port of the reference path) on this box's host cores, bounded sample

LK optical flow and the RANSAC fundamental matrix (src/Frame.cc:445-472) are not on the GPU in this round: the step
consumes precomputed previous-frame points and F ("what the step contains").
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
TH = 15.0                              # Tracking.cc:919-923 (RGB-D)


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


def make_track_inputs(kps, desc, counts, boxes, cap, point_cap, unique):
    """Per-frame inputs of the dyn-reject + match stage built from the extraction results (host side, untimed):
    previous-frame points (what LK would return), F, person boxes, u_right from a synthetic depth plane, and the last-frame
    map points = keypoints of the previous frame of the same stream back-projected with that depth."""
    from pysgs import synth
    import scenarios as S
    B = len(counts)
    cam = synth.TUM3
    depth = synth.depth_s1(W, H)
    rng = np.random.RandomState(1234)
    prev = np.zeros((B, cap, 2), np.float32); ur = np.full((B, cap), -1, np.float32)
    F = np.zeros((B, 9), np.float64); nb = np.ones(B, np.int32); have = np.ones(B, np.uint8)
    bx = np.zeros((B, 4, 4), np.float32); bx[:, 0] = boxes
    lxyz = np.zeros((B, point_cap, 3), np.float32); ldesc = np.zeros((B, point_cap, 32), np.uint8)
    lflags = np.zeros((B, point_cap), np.uint8); loct = np.zeros((B, point_cap), np.int32); lang = np.zeros((B, point_cap), np.float32)
    ln = np.zeros(B, np.int32)
    T = np.tile(np.eye(4, dtype=np.float32).reshape(1, 16), (B, 1))
    for f in range(B):
        n = counts[f]
        k = kps[f, :n]
        flow = np.array([2.0 + 0.5 * np.sin(0.3 * f), 1.0 * np.cos(0.2 * f)])          # image-plane pan of this frame
        t = flow / np.hypot(*flow)
        F[f] = np.array([[0, 0, t[1]], [0, 0, -t[0]], [-t[1], t[0], 0]]).reshape(9)     # epipolar lines parallel to the pan
        noise = rng.normal(0, 0.25, (n, 2))
        p = np.stack([k['x'], k['y']], 1) + flow + noise
        inbox = (k['x'] > boxes[f, 0]) & (k['x'] < boxes[f, 0] + boxes[f, 2]) & (k['y'] > boxes[f, 1]) & (k['y'] < boxes[f, 1] + boxes[f, 3])
        p[inbox] += np.array([-t[1], t[0]]) * rng.uniform(2.0, 6.0, (inbox.sum(), 1))  # the "person" moves off the epipolar lines
        prev[f, :n] = p
        z = depth[np.clip(k['y'].astype(np.int64), 0, H - 1), np.clip(k['x'].astype(np.int64), 0, W - 1)]
        ur[f, :n] = k['x'] - np.float32(cam['bf']) / z
        g = f - 1 if (f % unique) != 0 else f                                           # previous frame of the same stream
        m = min(counts[g], point_cap)
        kk = kps[g, :m]
        zz = depth[np.clip(kk['y'].astype(np.int64), 0, H - 1), np.clip(kk['x'].astype(np.int64), 0, W - 1)]
        lxyz[f, :m] = np.stack([(kk['x'] - cam['cx']) * zz / cam['fx'], (kk['y'] - cam['cy']) * zz / cam['fy'], zz], 1)
        ldesc[f, :m] = desc[g, :m]; loct[f, :m] = kk['octave']; lang[f, :m] = kk['angle']
        lflags[f, :m] = 1 | (2 * ((np.arange(m) % 5) != 0))                              # every 5th point is a temporal point (0 observations)
        ln[f] = m
    sf = S.scale_factors()
    return dict(prev=prev, ur=ur, F=F, boxes=bx, nb=nb, have=have, lxyz=lxyz, ldesc=ldesc, lflags=lflags, loct=loct, lang=lang, ln=ln, T=T, sf=sf)


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
        # "under load": samples above 60 % of the maximum seen, so idle gaps do not drag the median down
        load = [v for v in sm if v >= 0.6 * sm[-1]] or sm
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
def cpu_frames_per_s(frames, ti, nframes, cores):
    """The CPU oracle (port of the reference path: extract + dyn-reject + SearchByProjection) on `nframes` frames with `cores`
    worker threads (frames are independent; ctypes releases the GIL inside the C++ oracle)."""
    import oracle as O
    from concurrent.futures import ThreadPoolExecutor
    from pysgs import synth
    cam = synth.TUM3
    O.lib()

    def one(f):
        k, d = O.extract(frames[f])
        n = len(k)
        cur = np.stack([k['x'], k['y']], 1)
        _, keep, _, restored = O.dynreject(cur, ti['prev'][f, :n], ti['F'][f], ti['boxes'][f, :ti['nb'][f]], bool(ti['have'][f]), NFEAT)
        sel = np.arange(n) if restored else np.nonzero(keep)[0]
        fr = O.FrameArrays(k[sel], ti['ur'][f, :n][sel], d[sel], W, H, cam['fx'], cam['fy'], cam['cx'], cam['cy'], cam['bf'], ti['sf'])
        m = int(ti['ln'][f])
        nm, mp, nc = O.search_by_projection_last(fr, ti['T'][f].reshape(4, 4), ti['T'][f].reshape(4, 4), ti['lflags'][f, :m] & 1, ti['lxyz'][f, :m],
                                                 ti['ldesc'][f, :m], (ti['lflags'][f, :m] >> 1) & 1, ti['loct'][f, :m], ti['lang'][f, :m], TH)
        return n, len(sel), nm, mp, k, d, sel

    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        res = list(ex.map(one, range(nframes)))
    dt = time.perf_counter() - t0
    return nframes / dt, res


def run_reference(args):
    """--impl reference: the reference's own CPU path.  The reference cannot be compiled here (needs OpenCV/Eigen/ncnn/ROS,
    DESIGN.md), so this times the CPU oracle port with all host threads, on the same workload/config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    nb = per_step
    frames, boxes, unique = make_frames(nb, seed=2, unique=min(32, nb))
    import oracle as O
    O.lib()
    cap = NFEAT + 64
    kps = np.zeros((nb, cap), O.KP_DTYPE); desc = np.zeros((nb, cap, 32), np.uint8); counts = np.zeros(nb, np.int32)
    for f in range(nb):
        k, d = O.extract(frames[f]); counts[f] = len(k); kps[f, :len(k)] = k; desc[f, :len(k)] = d
    ti = make_track_inputs(kps, desc, counts, boxes, cap, cap, unique)
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
            'config': {'workload': 'S2 walking_xyz-shaped synthetic 640x480 stream, ORB 1000 features, extract+dyn-reject+SearchByProjection(th=15)',
                       'frames_per_step': per_step, 'note': 'CPU oracle port of the reference path (reference itself needs OpenCV/ROS: unbuildable here)'},
            'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': '%d frames per step x %d steps' % (per_step, args.steps)},
            'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=512, help='frames per GPU per step (512 x 307 KB = 157 MB of input > the 126 MB L2)')
    ap.add_argument('--cpu-sample', type=int, default=48, help='frames of the cpu_baseline sample')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
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

    # ---- workload + one-time set-up (untimed) -------------------------------------------------------------------------
    t_setup = time.time()
    frames, boxes, unique = make_frames(NB, seed=2 + rank)
    sf = S.scale_factors()
    cam = B.make_camera(W, H, synth.TUM3, sf)
    trk = B.Tracker(W, H, cam, NFEAT, 1.2, 8, 20, 7, max_batch=NB, point_cap=NFEAT + 64, max_boxes=4, device=local)
    cap, pcap = trk.cap, trk.point_cap
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_frames = pin((NB, H, W), torch.uint8); h_frames.numpy()[:] = frames
    h_kps = pin((NB, cap, 28), torch.uint8); h_desc = pin((NB, cap, 32), torch.uint8); h_n = pin((NB,), torch.int32)
    trk.extract(h_frames.data_ptr(), NB, W * H, W, h_kps.data_ptr(), h_desc.data_ptr(), h_n.data_ptr())
    kps0 = h_kps.numpy().reshape(NB, cap * 28).view(B.KP_DTYPE).reshape(NB, cap).copy(); desc0 = h_desc.numpy().copy(); n0 = h_n.numpy().copy()
    ti = make_track_inputs(kps0, desc0, n0, boxes, cap, pcap, unique)
    # the shared "map database" (last-frame descriptors + positions): rank 0's copy is broadcast once over NVLink
    keys_h = ['prev', 'ur', 'F', 'boxes', 'nb', 'have', 'lxyz', 'ldesc', 'lflags', 'loct', 'lang', 'ln', 'T']
    hp = {k: torch.from_numpy(np.ascontiguousarray(ti[k])).pin_memory() for k in keys_h}
    dv = {k: v.cuda(non_blocking=True) for k, v in hp.items()}
    d_frames = h_frames.cuda()
    bcast_ms = None
    if dist is not None:
        voc = torch.from_numpy(synth.descriptors_s5(1_081_000, 5)).cuda()   # ORBvoc-sized node-descriptor table (34.6 MB), SURVEY section 5
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); dist.broadcast(voc, 0); e1.record(); torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    h_out = dict(kps=pin((NB, cap, 28), torch.uint8), desc=pin((NB, cap, 32), torch.uint8), ur=pin((NB, cap), torch.float32), cnt=pin((NB,), torch.int32),
                 mp=pin((NB, cap), torch.int32), nm=pin((NB,), torch.int32))
    st = torch.cuda.Stream()
    ex_handle = B.lib().sgs_tracker_extractor
    ex_handle.restype = C.c_void_p
    exh = C.c_void_p(ex_handle(trk.h))
    torch.cuda.synchronize()
    log('[bench] rank %d set-up %.1fs: %d frames/step, mean %.0f keypoints/frame' % (rank, time.time() - t_setup, NB, n0.mean()))

    def track_ptrs(d):
        return [d[k].data_ptr() for k in ('prev', 'ur', 'F', 'boxes', 'nb', 'have', 'lxyz', 'ldesc', 'lflags', 'loct', 'lang', 'ln', 'T', 'T')]

    v = C.c_void_p

    def step_device():
        B.check(B.lib().sgs_tracker_extract_device(trk.h, v(d_frames.data_ptr()), NB, C.c_size_t(W * H), W, v(st.cuda_stream)))
        B.check(B.lib().sgs_tracker_track_device(trk.h, NB, *[v(p) for p in track_ptrs(dv)], C.c_float(TH), 0, 1, v(st.cuda_stream)))

    def step_host():
        # call 1: frames -> keypoints (what the host-side LK needs); descriptors stay on the device (desc = NULL)
        trk.extract(h_frames.data_ptr(), NB, W * H, W, h_kps.data_ptr(), 0, h_n.data_ptr())
        trk.track(NB, track_ptrs(hp), TH, 0, 1, [h_out['kps'].data_ptr(), h_out['desc'].data_ptr(), h_out['ur'].data_ptr(), h_out['cnt'].data_ptr(),
                                                 h_out['mp'].data_ptr(), h_out['nm'].data_ptr()])

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
            step_device()
    barrier()
    B.check(B.lib().sgs_extractor_set_profiling(exh, 1))
    sampler = ClockSampler(local); sampler.start(); time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    barrier()
    with torch.cuda.stream(st):
        ev[0].record(st)
        for i in range(args.steps):
            B.check(B.lib().sgs_tracker_extract_device(trk.h, v(d_frames.data_ptr()), NB, C.c_size_t(W * H), W, v(st.cuda_stream)))
            ev[3 * i + 1].record(st)
            B.check(B.lib().sgs_tracker_track_device(trk.h, NB, *[v(p) for p in track_ptrs(dv)], C.c_float(TH), 0, 1, v(st.cuda_stream)))
            ev[3 * i + 2].record(st)
            ev[3 * i + 3].record(st)
    barrier()
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[3 * args.steps]))
    extract_ms = sum(ev[3 * i].elapsed_time(ev[3 * i + 1]) for i in range(args.steps)) / args.steps
    track_ms = sum(ev[3 * i + 1].elapsed_time(ev[3 * i + 2]) for i in range(args.steps)) / args.steps
    clocks = sampler.stop()
    ms5 = (C.c_double * 5)(); ncalls = C.c_int()
    B.check(B.lib().sgs_extractor_stage_times(exh, ms5, C.byref(ncalls)))
    stage_ms = [ms5[i] / max(1, ncalls.value) for i in range(5)]
    B.check(B.lib().sgs_extractor_set_profiling(exh, 0))
    value = world * NB * args.steps / (total_ms * 1e-3)

    # results of the last device step: parity spot-check against the host path + counters for the byte accounting
    rp = [C.c_void_p() for _ in range(7)]
    B.check(B.lib().sgs_tracker_results_device(trk.h, *[C.byref(x) for x in rp]))
    step_host()   # also the e2e warm-up
    counts_after = h_out['cnt'].numpy().copy(); nmatch = h_out['nm'].numpy().copy()

    # ---- e2e leg: host buffers through the C ABI, copies inside the timed region ----------------------------------------
    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_host()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        h2d = h_frames.numel() + sum(hp[k].numel() * hp[k].element_size() for k in keys_h) + hp['T'].numel() * 4
        d2h = (h_kps.numel() + h_n.numel() * 4) + sum(t.numel() * t.element_size() for t in h_out.values())
        e2e = {'value': world * NB * args.steps / dt, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
               'ms_per_step': 1e3 * dt / args.steps, 'note': 'sgs_tracker_extract + sgs_tracker_track with pinned host buffers; the host-side LK/RANSAC between the two calls is not included (not on the GPU yet)'}

    # ---- roofline of the dominant kernel --------------------------------------------------------------------------------
    peaks, peak_kind = measured_peaks()
    names = ['pyramid(7 launches)', 'fast_cells_kernel', 'quadtree_kernel', 'blur(8 launches)', 'describe_kernel']
    dom = int(np.argmax(stage_ms))
    # FAST candidate total of one step (for the algorithmic bytes written by the FAST kernel)
    ncand_frame0 = 0
    for l in range(8):
        nn = C.c_int()
        B.lib().sgs_extractor_read_candidates(exh, 0, l, None, 0, C.byref(nn))   # count only (returns SGS_ERR_CAPACITY by design)
        ncand_frame0 += nn.value
    alg = {0: 1_569_878, 1: ALG_BYTES_FAST_READ + 4 * ncand_frame0, 2: 8 * ncand_frame0 + 4 * int(n0.mean()), 3: 1_901_064, 4: 749 * int(n0.mean()) + 544 * int(n0.mean()) + 60 * int(n0.mean())}
    dom_bytes = alg[dom] * NB
    achieved = dom_bytes / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'kernel': names[dom], 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'],
                'traffic': None, 'peak_kind': peak_kind + ' copy bandwidth (MEASURED_PEAKS.json)' if peak_kind == 'measured' else 'fallback 6650 GB/s',
                'algorithmic_bytes_per_launch': int(dom_bytes), 'kernel_ms': stage_ms[dom],
                'stage_ms': dict(zip(names, [round(x, 4) for x in stage_ms])), 'extract_ms': extract_ms, 'dynreject_match_ms': track_ms,
                'extract_alg_gbs': ALG_BYTES_EXTRACT * NB / (extract_ms * 1e-3) / 1e9,
                'extract_frac_of_hbm': ALG_BYTES_EXTRACT * NB / (extract_ms * 1e-3) / 1e9 / peaks['hbm_gbs']}

    # ---- CPU baseline on this box's host cores (rank 0 only, N=1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        cores = os.cpu_count() or 1
        ns = min(args.cpu_sample, NB)
        cpu_fps, res = cpu_frames_per_s(frames, ti, ns, cores)
        # parity of the sample: the oracle's result for these frames equals what the GPU returned through the host path
        kps_h = h_out['kps'].numpy().reshape(NB, cap * 28).view(B.KP_DTYPE).reshape(NB, cap)
        ok = True
        for f, (n, nsel, nm, mp, k, d, sel) in enumerate(res):
            ok &= int(counts_after[f]) == nsel and int(nmatch[f]) == nm and kps_h[f, :nsel].tobytes() == k[sel].tobytes()
            ok &= bool(np.array_equal(h_out['desc'].numpy()[f, :nsel], d[sel])) and bool(np.array_equal(h_out['mp'].numpy()[f, :nsel], mp))
        cpu = {'value': cpu_fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port', 'sample': '%d frames of the same batch, %d worker threads' % (ns, cores),
               'parity_with_gpu_on_sample': bool(ok)}
        if not ok:
            log('[bench] WARNING: GPU results differ from the oracle on the CPU sample')

    if rank == 0:
        line = {'metric': 'frames/sec ORB extract+match+dyn-reject 640x480', 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': warm, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8',
                'data': 'synthetic',
                'config': {'workload': 'S2 walking_xyz-shaped synthetic 640x480 stream (BASELINE configs[1]), ORB 1000 features / 8 levels / 1.2, extract + dyn-reject(geometry) + SearchByProjection(th=15)',
                           'frames_per_gpu_per_step': NB, 'l2_policy': 'inputs larger than L2: %d frames x 307200 B = %.0f MB per step (+ %.0f MB pyramid traffic)' % (NB, NB * 0.3072, NB * 0.95),
                           'sharding': 'independent streams per rank, no data-path collective; one untimed ncclBroadcast of the map/vocabulary table at start-up',
                           'mean_keypoints': float(n0.mean()), 'mean_after_dynreject': float(counts_after.mean()), 'mean_matches': float(nmatch.mean()),
                           'not_in_step': 'LK optical flow + RANSAC F (src/Frame.cc:445-472) and the detector: inputs precomputed'},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': 21 * args.steps, 'roofline': roofline, 'cpu_baseline': cpu}
        if bcast_ms is not None:
            line['config']['startup_broadcast_ms'] = bcast_ms
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
