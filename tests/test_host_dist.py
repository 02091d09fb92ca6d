"""world_size-2 gloo test of the frame-parallel plumbing on CPU (no GPU): two processes shard 5
frames, each renders its share with the CPU oracle, and the all-reduced loss sums must equal the
single-process sums."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ex4dgs_b200 import parallel, synth  # noqa: E402
from oracle import getters_oracle as GO  # noqa: E402

TIMESTAMPS = [3.0, 41.0, 137.0, 200.0, 288.0]


def _loss_of_frame(t):
    from tests import _util as U
    orc = U.oracle_module()
    sc = synth.make_scene(300, 100, 64, 48, sigma_px=3.0, seed=11)
    sc.timestamp = t
    inp = GO.flat_inputs(sc)
    rs = U.settings_for(orc, sc, "cpu")
    color = orc.GaussianRasterizer(rs)(means3D=inp["means3D"], means2D=torch.zeros_like(inp["means3D"]), dir3D=inp["dir3D"],
                                       opacities=inp["opacities"], shs=inp["shs"], scales=inp["scales"],
                                       rotations=inp["rotations"])[0]
    return color.abs().mean()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sums, mine = parallel.run_sharded(TIMESTAMPS, _loss_of_frame, rank, world)
    q.put((rank, [float(s) for s in sums], mine))
    dist.destroy_process_group()


def test_shard_frames():
    assert parallel.shard_frames(10, 0, 4) == [0, 4, 8]
    assert parallel.shard_frames(10, 3, 4) == [3, 7]
    assert sorted(sum((parallel.shard_frames(300, r, 8) for r in range(8)), [])) == list(range(300))
    with pytest.raises(ValueError):
        parallel.shard_frames(4, 4, 4)


def test_two_rank_gloo_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    single = [float(_loss_of_frame(t)) for t in TIMESTAMPS]
    expect = [single[0] + single[1], single[2] + single[3], single[4]]
    for rank, sums, mine in res:
        assert mine == parallel.shard_frames(len(TIMESTAMPS), rank, world)
        assert np.allclose(sums, expect, rtol=1e-6), (sums, expect)
