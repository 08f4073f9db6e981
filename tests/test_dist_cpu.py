"""N>1 plumbing of bench.py on CPU: world_size-2 gloo -- stream sharding is disjoint and deterministic, the start-up broadcast of the shared
descriptor table reaches every rank, and the max-over-ranks timing reduction works.  (The data path itself has no collective.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'sg-slam_b200'), os.path.join(ROOT, 'tests')]
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import bench
    from pysgs import synth
    frames, boxes, unique = bench.make_frames(6, 2 + rank, 640, 480, unique=3)
    table = torch.from_numpy(synth.descriptors_s5(4096, 5)) if rank == 0 else torch.zeros(4096, 32, dtype=torch.uint8)
    dist.broadcast(table, 0)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out[rank] = (int(frames.astype(np.uint64).sum()), int(table.numpy().astype(np.uint64).sum()), float(t.item()), frames.shape)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_broadcast():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert set(out.keys()) == {0, 1}
    s0, s1 = out[0], out[1]
    assert s0[3] == s1[3] == (6, 480, 640)
    assert s0[0] != s1[0]                      # different streams per rank
    assert s0[1] == s1[1] != 0                 # broadcast table identical everywhere
    assert s0[2] == s1[2] == 2.0               # max over ranks
    # determinism: the same rank seed gives the same shard
    sys.path[:0] = [ROOT, os.path.join(ROOT, 'sg-slam_b200'), os.path.join(ROOT, 'tests')]
    import bench
    f0, _, _ = bench.make_frames(6, 2, 640, 480, unique=3)
    assert int(f0.astype(np.uint64).sum()) == s0[0]


def test_reference_arm_prints_one_contract_line():
    """bench.py --impl reference: the CPU port of the path on the host cores, ONE JSON line on stdout with the contract's keys (no GPU involved)."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'frames/s' and d['value'] > 0 and d['n_gpus'] == 1 and d['higher_is_better'] is True
    assert d['cpu_baseline']['kind'] in ('port', 'reference') and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
