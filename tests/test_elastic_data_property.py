"""Property test of `ElasticDataAdaptor`: whatever sequence of cluster sizes a job goes through, the workers together consume the
global sample stream exactly once and in order — no sample is skipped or read twice across a resize (the reference's dataset adaptor has
the same contract: skip(offset).batch(b).shard(np, rank) with the offset carried over, v1/datasets/adaptor.py)."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

import kungfu_b200.elastic.datasets as ds


@settings(max_examples=120, deadline=None)
@given(sizes=st.lists(st.tuples(st.integers(1, 6), st.integers(1, 5)), min_size=1, max_size=6),  # (cluster size, steps at that size)
       batch=st.integers(1, 7), n=st.integers(20, 200))
def test_resizes_neither_skip_nor_repeat_samples(sizes, batch, n):
    data = torch.arange(n)
    labels = torch.arange(n) * 10
    state = {"np": 1, "rank": 0}
    old = ds.current_cluster_size, ds.current_rank
    ds.current_cluster_size = lambda: state["np"]
    ds.current_rank = lambda: state["rank"]
    try:
        consumed = []
        offset = 0  # what `ElasticState.trained_samples` carries across resizes
        for np_, steps in sizes:
            adaptors = [ds.ElasticDataAdaptor(data, labels, batch, offset=offset) for _ in range(np_)]  # survivors and joiners alike
            for _ in range(steps):
                step_batches = []
                for r, a in enumerate(adaptors):
                    state["np"], state["rank"] = np_, r
                    x, y = a.next_batch()
                    assert torch.equal(y, x * 10)  # labels travel with their samples
                    step_batches.append(x)
                consumed.append(torch.cat(step_batches))  # rank order = stream order inside a step
                offset += batch * np_
            assert all(a.offset == offset for a in adaptors)
        stream = torch.cat(consumed)
        want = torch.arange(len(stream)) % n
        assert torch.equal(stream, want)
    finally:
        ds.current_cluster_size, ds.current_rank = old


def test_reference_shaped_adaptor_shards_batches_round_robin():
    a = ds.ClusterDatasetAdaptor(batch_size=3, shard_count=2, shard_id=1)
    got = list(a(list(range(20))))
    assert got == [[3, 4, 5], [9, 10, 11], [15, 16, 17]]
    upd = a.create_update_offset()
    assert upd() == 6 and list(a(list(range(20))))[0] == [9, 10, 11]
    assert a.create_rewind()() == 0
