"""CPU, world_size 2 over gloo: the multi-GPU host logic of cnvkit_b200/shard.py (arm -> rank
assignment, local/global index mapping, count all_gather + padded gather) gives the same table as
the unsharded call.  The per-rank segmenter here is the CPU oracle -- the GPU call has the same
(log2, weight, arm_offsets) -> (arm, lo, hi, mean) contract."""
import os
import socket
import sys

import numpy as np
import pytest

from cnvkit_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_and_block_owner():
    own = shard.lpt_assign([9, 1, 1, 1, 4, 4], 2)
    loads = [sum(c for c, o in zip([9, 1, 1, 1, 4, 4], own) if o == r) for r in (0, 1)]
    assert sorted(loads) == [10, 10]
    assert shard.block_owner(10, 4).tolist() == [0, 0, 0, 1, 1, 2, 2, 2, 3, 3]
    assert shard.block_owner(0, 4).tolist() == []
    off = np.array([0, 10, 10, 110, 150])
    x = np.arange(150.0)
    owner = shard.arm_owner(off, 2)
    got = []
    for r in (0, 1):
        xs, ws, loff, mine = shard.take_arms(x, None, off, owner, r)
        assert ws is None and loff[-1] == len(xs)
        for k, a in enumerate(mine):
            np.testing.assert_array_equal(xs[loff[k]:loff[k + 1]], x[off[a]:off[a + 1]])
            got.append(a)
    assert sorted(got) == [0, 1, 2, 3]


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from oracle import cbs as OC

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        sizes = [300, 40, 0, 1200, 7, 650, 2]
        xs = []
        for n in sizes:
            x = rng.normal(0, 0.2, n)
            if n > 100:
                x[n // 3: n // 2] += 0.8
            xs.append(x)
        x = np.concatenate(xs)
        w = rng.uniform(0.5, 1, len(x))
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        P = OC.make_params(alpha=1e-4)

        def seg(lx, lw, loff):
            return OC.segment(lx, lw, loff, P, nthreads=1)

        full = shard.segment_arms_sharded(x, w, off, seg, dst=0)
        # a rank with no rows at all must still take part in the gather
        empty = shard.gather_tables([np.zeros(0, np.int64), np.zeros(0)] if rank == 1 else
                                    [np.arange(3, dtype=np.int64), np.ones(3)], dst=0)
        capped = shard.gather_tables([np.arange(rank + 2, dtype=np.int64), np.full(rank + 2, 0.5 + rank)], dst=0,
                                     capacity=8)
        # batched gather: several small tables appended per rank, ONE collective per flush, reusable
        tg = shard.TableGatherer(ncols=3, capacity=16)
        for rnd in range(2):
            for step in range(2 + rank):
                k = step + 1
                tg.append([np.full(k, rank, dtype=np.int64), np.arange(k) + 10 * step, np.full(k, 0.25 * rnd)])
            tab = tg.flush(dst=0)
            if rank == 0:
                assert tab.shape == (3 + 6, 3)
                assert tab[:, 0].tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 1]
                assert tab[:, 1].tolist() == [0, 10, 11, 0, 10, 11, 20, 21, 22]
                assert (tab[:, 2] == 0.25 * rnd).all()
            else:
                assert tab is None
        with pytest.raises(ValueError):
            tg.append([np.zeros(17), np.zeros(17), np.zeros(17)])
        if rank == 0:
            assert capped[0].tolist() == [0, 1, 0, 1, 2] and capped[1].tolist() == [0.5, 0.5, 1.5, 1.5, 1.5]
        else:
            assert capped is None
        if rank == 0:
            ref = OC.segment(x, w, off, P, nthreads=1)
            for a, b in zip(full, ref[:4]):
                np.testing.assert_array_equal(a, b)
            assert full[0].dtype == np.int64 and full[3].dtype == np.float64
            assert empty[0].tolist() == [0, 1, 2] and empty[0].dtype == np.int64
            open(out_path, "w").write("ok %d" % len(full[0]))
        else:
            assert full is None and empty is None
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_segmentation_world2(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = tmp_path / "ok.txt"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    assert out.read_text().startswith("ok")
