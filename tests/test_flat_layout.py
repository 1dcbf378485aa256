"""Property tests of the flat gradient / parameter layout and its bucket plan (kungfu_b200/parallel/flat.py): whatever the tensor sizes,
bucket size and tail size, every entry lies inside exactly one bucket, buckets tile the region without gaps or overlaps, every tensor and
every bucket starts on a 128-byte boundary, shards partition a bucket, and the layout digest only depends on what the ranks must agree on."""
import torch
from hypothesis import given, settings
from hypothesis import strategies as st

from kungfu_b200.parallel.flat import ALIGN_BYTES, FlatLayout

sizes = st.lists(st.integers(min_value=1, max_value=70_000), min_size=1, max_size=40)
dtypes = st.sampled_from([torch.float32, torch.bfloat16, torch.float16])


def params_of(ns, dtype):
    return [(f"p{i}", torch.nn.Parameter(torch.zeros(n, dtype=dtype))) for i, n in enumerate(ns)]


@settings(max_examples=150, deadline=None)
@given(ns=sizes, dtype=dtypes, bucket_kb=st.integers(min_value=1, max_value=512), tail_kb=st.integers(min_value=0, max_value=64),
       reverse=st.booleans())
def test_buckets_tile_the_region(ns, dtype, bucket_kb, tail_kb, reverse):
    lay = FlatLayout(params_of(ns, dtype), dtype, bucket_kb << 10, reverse=reverse, tail_bytes=tail_kb << 10)
    align = ALIGN_BYTES // lay.esize
    assert len(lay.entries) == len(ns) and sum(e.numel for e in lay.entries) == sum(ns)
    order = [e.name for e in lay.entries]
    assert order == [f"p{i}" for i in (range(len(ns) - 1, -1, -1) if reverse else range(len(ns)))]
    # entries: aligned, ordered, non-overlapping
    end = 0
    for e in lay.entries:
        assert e.offset % align == 0 and e.offset >= end
        end = e.offset + e.numel
    # buckets: aligned, contiguous, cover [0, total), indices consecutive
    pos = 0
    for i, b in enumerate(lay.buckets):
        assert b.index == i and b.start == pos and b.start % align == 0 and b.numel % align == 0 and b.numel > 0
        pos = b.start + b.numel
    assert pos == lay.total and lay.nbytes == lay.total * lay.esize
    # every entry belongs to exactly one bucket, and lies inside it
    seen = []
    for b in lay.buckets:
        assert b.entries, "no empty buckets"
        for ei in b.entries:
            e = lay.entries[ei]
            assert e.bucket == b.index and b.start <= e.offset and e.offset + e.numel <= b.start + b.numel
            seen.append(ei)
    assert sorted(seen) == list(range(len(ns)))
    # bucket order follows entry order (what backward produces first is reduced first)
    assert [ei for b in lay.buckets for ei in b.entries] == list(range(len(ns)))


@settings(max_examples=100, deadline=None)
@given(ns=sizes, dtype=dtypes, bucket_kb=st.integers(min_value=1, max_value=256), nranks=st.integers(min_value=1, max_value=8))
def test_shards_partition_every_bucket(ns, dtype, bucket_kb, nranks):
    lay = FlatLayout(params_of(ns, dtype), dtype, bucket_kb << 10, tail_bytes=4 << 10)
    lanes = 16 // lay.esize
    for b in lay.buckets:
        per = lay.shard_elems(b, nranks)
        assert per % lanes == 0 and per * nranks >= b.numel
        pos = b.start
        for r in range(nranks):
            lo, hi = lay.shard_range(b, nranks, r)
            assert lo == pos and lo <= hi <= b.start + b.numel and hi - lo <= per
            pos = hi
        assert pos == b.start + b.numel


@settings(max_examples=50, deadline=None)
@given(ns=sizes, bucket_kb=st.integers(min_value=1, max_value=256))
def test_tail_bucket_is_small_and_last(ns, bucket_kb):
    tail = 8 << 10
    with_tail = FlatLayout(params_of(ns, torch.bfloat16), torch.bfloat16, bucket_kb << 10, tail_bytes=tail)
    without = FlatLayout(params_of(ns, torch.bfloat16), torch.bfloat16, bucket_kb << 10, tail_bytes=0)
    assert with_tail.total == without.total and [e.offset for e in with_tail.entries] == [e.offset for e in without.entries]
    assert len(with_tail.buckets) in (len(without.buckets), len(without.buckets) + 1)
    if len(with_tail.buckets) == len(without.buckets) + 1:
        last = with_tail.buckets[-1]
        payload = sum(with_tail.entries[i].numel for i in last.entries)
        assert payload * 2 <= tail  # only tensors that fit the tail budget were moved
        assert last.start + last.numel == with_tail.total and with_tail.buckets[-2].start == without.buckets[-1].start
    assert with_tail.digest() != without.digest() or len(with_tail.buckets) == len(without.buckets)


def test_digest_depends_on_names_sizes_and_plan():
    a = FlatLayout(params_of([10, 20, 30], torch.float32), torch.float32, 1 << 10)
    b = FlatLayout(params_of([10, 20, 30], torch.float32), torch.float32, 1 << 10)
    c = FlatLayout(params_of([10, 20, 31], torch.float32), torch.float32, 1 << 10)
    d = FlatLayout(params_of([10, 20, 30], torch.float32), torch.float32, 64)
    assert a.digest() == b.digest() and a.digest() != c.digest() and a.digest() != d.digest()
