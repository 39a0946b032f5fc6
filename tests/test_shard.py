"""N > 1 host logic on two gloo ranks: disjoint stream assignment, every rank
parses its own streams through the plugin front end (parse only, no device), and the job
throughput is units of all ranks over the slowest rank's time."""
import os
import socket
import sys

import pytest

import streams
import util

torch = pytest.importorskip("torch")
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, util.ROOT)
    sys.path.insert(0, os.path.join(util.ROOT, "tests"))
    import torch.distributed as dist
    import m2dec_b200
    from m2dec_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seeds = shard.rank_seeds(rank, 3)
        pictures = 0
        for s in seeds:
            data = streams.generate("h264", "mbw=6 mbh=4 frames=5 seed=%d" % s)
            seen = []
            m2dec_b200.decode_file("h264", data, tap=lambda dec, codec, pic: seen.append(codec) or 0, device_off=True)
            pictures += len(seen)
        dist.barrier()
        seconds = 1.0 + rank          # rank 1 is the slow one
        rate = shard.whole_job_rate(pictures, seconds, dist)
        worst = shard.max_over_ranks(seconds, dist)
        with open(os.path.join(out_dir, "r%d.txt" % rank), "w") as f:
            f.write("%s|%d|%r|%r\n" % (",".join(map(str, seeds)), pictures, rate, worst))
    finally:
        dist.destroy_process_group()


def test_two_ranks_partition_and_aggregate(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rows = [open(tmp_path / ("r%d.txt" % r)).read().strip().split("|") for r in range(world)]
    seeds = [set(r[0].split(",")) for r in rows]
    assert not (seeds[0] & seeds[1]) and all(len(s) == 3 for s in seeds)
    assert [int(r[1]) for r in rows] == [15, 15]            # 3 streams x 5 pictures on every rank
    # both ranks agree on the aggregate: 30 pictures / the slowest rank's 2.0 s
    assert [float(r[2]) for r in rows] == [15.0, 15.0]
    assert [float(r[3]) for r in rows] == [2.0, 2.0]


def test_single_process_is_identity():
    from m2dec_b200 import shard
    assert shard.whole_job_rate(10, 2.0) == 5.0
    assert shard.rank_seeds(2, 2) == [2100, 2101]
    with pytest.raises(ValueError):
        shard.rank_seeds(0, 5000)
