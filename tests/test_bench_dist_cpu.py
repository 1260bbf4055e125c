"""CPU, world_size 2 over gloo: the host-side logic of the N > 1 path (utterance sharding, max-over-ranks timing,
the reference arm running on rank 0 only)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from eabnet_b200.shard import shard_range, max_over_ranks
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
b, e = shard_range(2048, r, w)
mine = torch.zeros(2048, dtype=torch.int32); mine[b:e] = 1
dist.all_reduce(mine)
ms = max_over_ranks([10.0 + r, 5.0 - r])
if r == 0:
    print(json.dumps({"covered_once": bool((mine == 1).all()), "ms": ms, "world": w}))
dist.destroy_process_group()
'''


def _torchrun(args, timeout=300):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29531"] + args
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


def test_shard_range_properties():
    from eabnet_b200.shard import shard_range
    for n in (0, 1, 7, 64, 2048, 2049):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [e - b for b, e in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = _torchrun([str(script)])
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line == {"covered_once": True, "ms": [11.0, 5.0], "world": 2}


def test_reference_arm_prints_once_under_torchrun():
    r = _torchrun([os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                   "--ref-batch", "1", "--seconds", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and lines[0]["impl"] == "reference" and lines[0]["value"] > 0
    assert lines[0]["cpu_baseline"]["kind"] == "port" and lines[0]["n_gpus"] == 2
