"""world_size-2 (and 3) gloo tests of the sharding host logic on CPU: the partition, the all-gather of verdict bitmaps
and digests, and shard-count independence of the results.  The per-shard compute is an oracle-backed stand-in here
(this file is test code); on GPUs it is pbh_b200.sharding.gpu_compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _oracle_compute(first, count):
    import oracle as O
    w, r, c, u, _ = O.generate_inputs(count, first_index=first, seed=0xB200, dist=1)
    proof, status = O.prove_batch(w, r, c)
    res = O.verify_batch(proof, c, u, want_gt=False)
    d = O.digest(proof, first_index=first)
    d = d - 2**64 if d >= 2**63 else d
    return torch.from_numpy(O.pack_verdicts(res)), torch.tensor([d], dtype=torch.int64)


def _worker(rank, world, port, n_total, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "plonk-by-fingers_b200", "python"), os.path.join(root, "oracle")]
    from pbh_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bitmap, digs, total = sharding.run_sharded(n_total, rank, world, _oracle_compute)
    torch.save({"bitmap": bitmap, "digs": digs, "total": total}, os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_ranges_cover_and_align():
    from pbh_b200 import sharding
    for n in (0, 1, 7, 8, 9, 1000, 1 << 20, (1 << 28) + 5):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = sharding.shard_range(n, r, world)
                assert lo == prev and lo <= hi <= n and (lo % 8 == 0 or lo == n)
                prev = hi
            assert prev == n


@pytest.mark.parametrize("world", (2, 3))
def test_gather_is_shard_count_independent(tmp_path, world):
    from pbh_b200 import sharding
    n_total = 3001     # not a multiple of 8 or of the world size
    single_bitmap, single_digs, single_total = sharding.run_sharded(n_total, 0, 1, _oracle_compute)
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    for o in outs:
        assert torch.equal(o["bitmap"], single_bitmap)        # same bytes on every rank, whatever the shard count
        assert o["total"] == single_total
        assert o["digs"].numel() == world
    assert int(np.unpackbits(single_bitmap.numpy(), bitorder="little")[:n_total].sum()) > 0
