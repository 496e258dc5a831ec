"""world_size-2 gloo tests (CPU) of the host-side logic of the multi-GPU path: task/marker partition, the merge
order of changed markers, and that summed per-rank statistics reproduce the single-process chain of the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        from hydra_b200 import partition
        from helpers import random_bed, reference_lists, simulate_y
        M, T, N, SR, K = 203, 6, 400, 4, 4
        lay = partition.rank_layout(M, T, world, rank)
        # 1. the ranks tile the markers without gaps, and agree on lmax
        t = torch.tensor([lay["m_start"], lay["m_local"], lay["lmax"]])
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        pos = 0
        for v in allv:
            assert int(v[0]) == pos and int(v[2]) == int(allv[0][2])
            pos += int(v[1])
        assert pos == M
        s, l = oracle.define_blocks(M, T)
        ps, pl = partition.define_blocks(M, T)
        assert s.tolist() == ps.tolist() and l.tolist() == pl.tolist()
        # 2. global window positions: every (step, task) slot is owned exactly once
        tl = lay["tasks_local"]
        mine = torch.tensor([partition.global_position(p, tl, T, lay["task_first"]) for p in range(SR * tl)])
        allp = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allp, mine)
        assert sorted(torch.cat(allp).tolist()) == list(range(SR * T))
        # 3. merged order of changed markers is the same on every rank and sorted by (step, task)
        rng = np.random.default_rng(rank)
        local = [(int(mine[i]), rank * 1000 + i) for i in sorted(rng.choice(SR * tl, 5, replace=False))]
        gathered = [None] * world
        dist.all_gather_object(gathered, local)
        merged = partition.merge_changed(gathered)
        assert [e[0] for e in merged] == sorted(e[0] for e in merged)
        chk = [None] * world
        dist.all_gather_object(chk, merged)
        assert all(c == chk[0] for c in chk)
        # 4. per-rank group statistics summed with all_reduce equal the oracle's whole-chain statistics
        rng = np.random.default_rng(0)
        bed, g = random_bed(rng, M, N)
        sp = reference_lists(bed, N)
        y = simulate_y(rng, g)
        tape = oracle.TapeMaker(5, T, M).make(2)
        ref = oracle.brr_chain(N, M, T, K, 1, SR, 2, sp, y, np.zeros(M, np.int32), np.array([[0, 0.001, 0.01, 0.1]]), tape, np.array([0.5]),
                               hyper_seed=3)
        sl = slice(lay["m_start"], lay["m_start"] + lay["m_local"])
        part = torch.tensor([float((ref["beta"][1][sl] ** 2).sum())] + [float((ref["comp"][1][sl] == k).sum()) for k in range(K)], dtype=torch.float64)
        dist.all_reduce(part)
        np.testing.assert_allclose(part[0].item(), ref["bsq"][1][0], rtol=1e-12)
        assert [int(x) for x in part[1:]] == ref["cass"][1][0].tolist()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_layout_and_merge_with_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
