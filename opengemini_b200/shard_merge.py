"""Cross-shard merge of per-GPU dense interval records (configs[3]).

One shard per GPU, no collective on the data path; this is the single exchange step at the end, the GPU counterpart of the
reference's partial-aggregate merge between store and sql nodes (engine/executor/agg_transform.go:248-304 for the reducers,
lib/record/reccord_functions.go:482-494 for the value/time tie-break).

    sum / count           all_reduce(SUM) on values, all_reduce(MAX) on validity
    min/max/first/last    all_gather of (value, valid, time) then an ordered fold rank 0..N-1 on every rank, through
                          `fold(remote_cols)` — on GPUs that is og_query_merge_dense (k_merge_dense keeps the reference's
                          tie-breaks); tests inject their own fold to exercise this host logic over gloo.

Float sums over ranks are reduced in the collective's order, not the reference's arrival order: covered by the 1e-9 relative
bound (north_star).  Everything else is bit-exact.
"""
from . import _lib as L

ADDITIVE = (L.AGG_SUM, L.AGG_COUNT)


def shard_seed(base_seed, rank):
    """Distinct synthetic shard per rank (configs[3]: '8 shards x 10k series')."""
    return int(base_seed) + 1000003 * int(rank)


def series_range_for_rank(n_series, rank, world):
    """If there are fewer shards than GPUs a shard is split by contiguous series range (a chunk = one series' pages,
    engine/immutable/chunkdata_builder_ts.go:36-82).  Returns [begin, end)."""
    base, rem = divmod(int(n_series), int(world))
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def cross_shard_merge(torch, dist, cols, world, fold):
    """cols: list of dict(values, valid, times|None, func) of torch tensors (this rank's partial, updated in place).
    fold(remote): called once per remote rank r=1..world-1, in rank order, with `remote` = list aligned with cols holding
    (values, valid, times|None) for selector columns and None for additive ones; it must fold them into this rank's cols
    with the reference tie-break."""
    if world == 1:
        return
    sel = []
    for i, c in enumerate(cols):
        if c["func"] in ADDITIVE:
            dist.all_reduce(c["values"], op=dist.ReduceOp.SUM)
            dist.all_reduce(c["valid"], op=dist.ReduceOp.MAX)
        else:
            sel.append(i)
    if not sel:
        return
    gathered = {}
    for i in sel:
        c = cols[i]
        gv = [torch.empty_like(c["values"]) for _ in range(world)]
        gk = [torch.empty_like(c["valid"]) for _ in range(world)]
        dist.all_gather(gv, c["values"])
        dist.all_gather(gk, c["valid"])
        gt = None
        if c["times"] is not None:
            gt = [torch.empty_like(c["times"]) for _ in range(world)]
            dist.all_gather(gt, c["times"])
        gathered[i] = (gv, gk, gt)
    # every rank folds ranks 0..world-1 in rank order, so all ranks end with identical bits: start from rank 0's partial
    for i in sel:
        gv, gk, gt = gathered[i]
        cols[i]["values"].copy_(gv[0])
        cols[i]["valid"].copy_(gk[0])
        if gt is not None:
            cols[i]["times"].copy_(gt[0])
    for r in range(1, world):
        remote = [None] * len(cols)
        for i in sel:
            gv, gk, gt = gathered[i]
            remote[i] = (gv[r], gk[r], gt[r] if gt is not None else None)
        fold(remote)


def gpu_fold(torch, q):
    """fold() for the product path: og_query_merge_dense on this query's dense record."""
    import ctypes as C

    def fold(remote):
        torch.cuda.synchronize()
        dv = q.dense_view()
        cols = (L.DenseCol * dv.n_cols)()
        n = dv.n_groups * dv.n_buckets
        zero_ok = None
        for i in range(dv.n_cols):
            cols[i] = dv.cols[i]
            if remote[i] is None:  # additive columns were all-reduced already: present an all-invalid partial
                if zero_ok is None:
                    zero_ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
                cols[i].valid = zero_ok.data_ptr()
            else:
                v, k, t = remote[i]
                cols[i].values, cols[i].valid = v.data_ptr(), k.data_ptr()
                cols[i].times = t.data_ptr() if t is not None else None
        other = L.DenseView(dv.n_groups, dv.n_buckets, dv.start, dv.interval, dv.n_cols, cols, None)
        L.check(L.lib().og_query_merge_dense(q.h, C.byref(other)), "og_query_merge_dense")
        torch.cuda.synchronize()
    return fold
