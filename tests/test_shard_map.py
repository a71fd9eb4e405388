"""Property tests (hypothesis) of the rendezvous-hashed ownership map for replicated virtual shards."""

import pytest

pytest.importorskip("hypothesis")

from hypothesis import given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402

from torchft_b200.parallel.shard_map import ShardMap  # noqa: E402

ids = st.lists(st.integers(0, 40).map(lambda i: f"replica_{i}"), min_size=1, max_size=12, unique=True)


@settings(max_examples=60, deadline=None)
@given(members=ids, k=st.integers(1, 3))
def test_every_shard_has_k_distinct_owners_and_order_is_irrelevant(members, k):
    sm = ShardMap(num_shards=64, replication=k)
    for s in range(sm.num_shards):
        own = sm.owners(s, members)
        assert len(own) == min(k, len(members)) and len(set(own)) == len(own) and set(own) <= set(members)
        assert own == sm.owners(s, list(reversed(members)))
    total = sum(len(v) for v in sm.assignment(members).values())
    assert total == sm.num_shards * min(k, len(members))
    assert sum(len(v) for v in sm.primaries(members).values()) == sm.num_shards


@settings(max_examples=60, deadline=None)
@given(members=ids.filter(lambda m: len(m) >= 2), k=st.integers(1, 3), data=st.data())
def test_leaving_replica_moves_only_its_own_shards(members, k, data):
    sm = ShardMap(num_shards=64, replication=k)
    gone = data.draw(st.sampled_from(members))
    rest = [m for m in members if m != gone]
    t = sm.plan_transition(members, rest)
    held_by_gone = set(sm.assignment(members)[gone])
    assert {s for s, _, _ in t.copies} <= held_by_gone  # nobody else's shard moves
    for s, src, dst in t.copies:
        assert src in sm.owners(s, members) and src != gone and dst in sm.owners(s, rest)
    # a shard is lost exactly when the leaver was its ONLY holder
    assert set(t.lost) == {s for s in held_by_gone if len(sm.owners(s, members)) == 1}
    # after the copies, every owner of every surviving shard holds it
    receives = {(s, dst) for s, _, dst in t.copies}
    for s in range(sm.num_shards):
        if s in t.lost:
            continue
        for dst in sm.owners(s, rest):
            assert dst in sm.owners(s, members) or (s, dst) in receives


@settings(max_examples=60, deadline=None)
@given(members=ids, k=st.integers(1, 3), new=st.integers(41, 60).map(lambda i: f"replica_{i}"))
def test_joining_replica_only_takes_over(members, k, new):
    sm = ShardMap(num_shards=64, replication=k)
    t = sm.plan_transition(members, members + [new])
    assert not t.lost
    if len(members) >= k:
        assert all(dst == new for _, _, dst in t.copies)  # existing owners never trade shards among themselves
    else:
        assert all(dst == new or dst in members for _, _, dst in t.copies)
    for s, src, _ in t.copies:
        assert src in sm.owners(s, members)


def test_double_failure_loses_only_shards_held_by_both_and_load_is_balanced():
    sm = ShardMap(num_shards=512, replication=2)
    members = [f"replica_{i}" for i in range(8)]
    load = [len(v) for v in sm.assignment(members).values()]
    assert sum(load) == 1024 and min(load) > 0.7 * 128 and max(load) < 1.3 * 128  # ~128 each
    prim = [len(v) for v in sm.primaries(members).values()]
    assert sum(prim) == 512 and min(prim) > 0.6 * 64 and max(prim) < 1.4 * 64
    a, b = members[2], members[5]
    rest = [m for m in members if m not in (a, b)]
    t = sm.plan_transition(members, rest)
    both = {s for s in range(512) if set(sm.owners(s, members)) == {a, b}}
    assert set(t.lost) == both and 0 < len(both) < 40  # expected 512 / C(8,2) ~ 18
    # first quorum ever: nothing to copy, nothing lost
    t0 = sm.plan_transition([], members)
    assert not t0.copies and not t0.lost and t0.kept == 0
