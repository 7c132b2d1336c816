"""N>1 host logic on CPU: world_size-2 gloo run of opengemini_b200.shard_merge (the exchange step of configs[3]).

The product fold is og_query_merge_dense on the GPU (covered by test_gpu_parity / bench --gpus 2); here the fold is a numpy
restatement of the reference tie-break (lib/record/reccord_functions.go:482-494 min/max, :640-700 first/last), so the
collective plumbing — which columns are all-reduced, gather order, rank-ordered fold, identical bits on every rank — is
what is under test.
"""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partial(rank, n, funcs):
    """Deterministic per-rank partial: some cells invalid, ties across ranks on purpose."""
    rng = np.random.default_rng(100 + rank)
    cols = []
    for f in funcs:
        valid = (rng.random(n) < 0.7).astype(np.uint8)
        if f == 1:  # count
            vals = rng.integers(1, 50, n).astype(np.int64) * valid
            times = None
        elif f == 2:  # sum (float bits)
            vals = (rng.random(n) * 100 * valid).view(np.int64)
            times = None
        else:  # selectors: coarse values so that ranks tie
            vals = np.floor(rng.random(n) * 4).view(np.int64)
            times = rng.integers(0, 6, n).astype(np.int64)
        cols.append(dict(values=vals.copy(), valid=valid, times=times, func=f))
    return cols


def _fold_numpy(cols, remote):
    """Reference rule for one remote partial (cross-series update of a dense interval record)."""
    for c, r in zip(cols, remote):
        if r is None:
            continue
        v, k, t = r
        f = c["func"]
        a = c["values"].view(np.float64)
        b = v.view(np.float64)
        for i in range(len(k)):
            if not k[i]:
                continue
            if not c["valid"][i]:
                take = True
            elif f == 3:  # min: smaller value, tie -> earlier time
                take = b[i] < a[i] or (b[i] == a[i] and t[i] < c["times"][i])
            elif f == 4:  # max
                take = b[i] > a[i] or (b[i] == a[i] and t[i] < c["times"][i])
            elif f == 5:  # first: earlier time, tie -> larger value
                take = t[i] < c["times"][i] or (t[i] == c["times"][i] and b[i] > a[i])
            else:  # last: later time, tie -> larger value
                take = t[i] > c["times"][i] or (t[i] == c["times"][i] and b[i] > a[i])
            if take:
                c["values"][i], c["times"][i], c["valid"][i] = v[i], t[i], 1


def _worker(rank, world, port, funcs, n, out):
    import torch
    import torch.distributed as dist
    from opengemini_b200 import shard_merge

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = _partial(rank, n, funcs)
        cols = []
        for c in mine:
            is_sum = c["func"] == 2
            vals = torch.from_numpy(c["values"].view(np.float64) if is_sum else c["values"])
            cols.append(dict(values=vals, valid=torch.from_numpy(c["valid"]),
                             times=None if c["times"] is None else torch.from_numpy(c["times"]), func=c["func"]))

        def fold(remote):
            rem = [None if r is None else (r[0].numpy(), r[1].numpy(), r[2].numpy()) for r in remote]
            view = [dict(values=c["values"].numpy().view(np.int64), valid=c["valid"].numpy(),
                         times=None if c["times"] is None else c["times"].numpy(), func=c["func"]) for c in cols]
            _fold_numpy(view, rem)

        shard_merge.cross_shard_merge(torch, dist, cols, world, fold)
        res = [(c["values"].numpy().view(np.int64).copy(), c["valid"].numpy().copy(),
                None if c["times"] is None else c["times"].numpy().copy()) for c in cols]
        out.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_cross_shard_merge_world2_gloo():
    import torch.multiprocessing as mp

    funcs = [2, 1, 4, 3, 5, 6]  # sum, count, max, min, first, last
    n, world = 257, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, funcs, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # expected: sequential fold rank 0 then rank 1
    parts = [_partial(r, n, funcs) for r in range(world)]
    exp = parts[0]
    for k, c in enumerate(exp):
        o = parts[1][k]
        if c["func"] == 1:
            c["values"] = c["values"] + o["values"]
            c["valid"] = np.maximum(c["valid"], o["valid"])
        elif c["func"] == 2:
            c["values"] = (c["values"].view(np.float64) + o["values"].view(np.float64)).view(np.int64)
            c["valid"] = np.maximum(c["valid"], o["valid"])
    _fold_numpy(exp, [None if c["func"] in (1, 2) else (parts[1][k]["values"], parts[1][k]["valid"], parts[1][k]["times"])
                      for k, c in enumerate(exp)])
    for r in range(world):
        for k, c in enumerate(exp):
            v, ok, t = got[r][k]
            assert np.array_equal(ok, c["valid"]), (r, k)
            m = ok.astype(bool)
            assert np.array_equal(v[m], c["values"][m]), (r, k)  # bit-exact, identical on both ranks
            if c["times"] is not None:
                assert np.array_equal(t[m], c["times"][m]), (r, k)


def test_rank_partition_helpers():
    from opengemini_b200 import shard_merge
    assert len({shard_merge.shard_seed(1000, r) for r in range(8)}) == 8
    for n, w in [(10000, 8), (10, 4), (7, 8), (1, 2)]:
        spans = [shard_merge.series_range_for_rank(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
