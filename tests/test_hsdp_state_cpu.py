"""Heal payload of an FSDP2-sharded replica (``parallel/hsdp.py:fsdp_local_state``) on CPU: FSDP2 over a one-rank gloo
mesh is enough to get real DTensor parameters and optimizer state. A trained replica's payload installed on a fresh one
must make the two indistinguishable: same weights, same moments, same next step (reference heal semantics:
manager.py:700-716 -- rank i of the joiner loads what rank i of a healthy group serves)."""
import os

import pytest
import torch
import torch.distributed as dist
from torch import nn

from torchft_b200.parallel.hsdp import fsdp_local_state, load_fsdp_local_state


@pytest.fixture(scope="module")
def gloo_world():
    if dist.is_initialized():
        pytest.skip("a default process group already exists in this interpreter")
    keys = ("MASTER_ADDR", "MASTER_PORT", "RANK", "WORLD_SIZE")
    saved = {k: os.environ.get(k) for k in keys}
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29917", RANK="0", WORLD_SIZE="1")
    dist.init_process_group("gloo")
    try:
        yield
    finally:
        dist.destroy_process_group()
        for k, v in saved.items():  # later tests spawn torchrun workers: do not leak a rendezvous into their environment
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _replica(seed: int):
    from torch.distributed.device_mesh import init_device_mesh
    from torch.distributed.fsdp import fully_shard

    torch.manual_seed(seed)
    m = nn.Sequential(nn.Linear(8, 16), nn.GELU(), nn.Linear(16, 4))
    mesh = init_device_mesh("cpu", (1,))
    for layer in m:
        if isinstance(layer, nn.Linear):
            fully_shard(layer, mesh=mesh)
    fully_shard(m, mesh=mesh)
    return m, torch.optim.AdamW(m.parameters(), lr=1e-2, weight_decay=0.1)


def _step(m, opt, x):
    opt.zero_grad(set_to_none=True)
    m(x).pow(2).mean().backward()
    opt.step()


def test_payload_of_a_trained_replica_makes_a_fresh_one_identical(gloo_world):
    torch.manual_seed(0)
    xs = [torch.randn(5, 8) for _ in range(4)]
    a, oa = _replica(1)
    for x in xs[:3]:
        _step(a, oa, x)
    b, ob = _replica(2)  # different init, no optimizer state yet
    payload = fsdp_local_state(a, oa)
    assert set(payload) == {"params", "optim"} and set(payload["params"]) == {n for n, _ in a.named_parameters()}
    # the payload aliases the live storage (an in-place transport writes straight into it)
    p0 = next(a.parameters())
    assert payload["params"]["0.weight"].data_ptr() == p0.detach().to_local().data_ptr()
    load_fsdp_local_state(b, ob, {k: {n: (v.clone() if isinstance(v, torch.Tensor) else {kk: vv.clone() for kk, vv in v.items()})
                                      for n, v in d.items()} for k, d in payload.items()})
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa.full_tensor(), pb.full_tensor()), n
        sa, sb = oa.state[pa], ob.state[pb]
        assert set(sa) == set(sb)
        for k in sa:
            va, vb = sa[k], sb[k]
            assert type(va) is type(vb), (n, k)  # DTensor moments stay DTensors with the parameter's sharding
            assert torch.equal(va.full_tensor() if hasattr(va, "full_tensor") else va, vb.full_tensor() if hasattr(vb, "full_tensor") else vb)
    _step(a, oa, xs[3])
    _step(b, ob, xs[3])
    for pa, pb in zip(a.parameters(), b.parameters()):
        assert torch.equal(pa.full_tensor(), pb.full_tensor())


def test_in_place_targets_are_left_alone_and_shape_mismatch_is_loud(gloo_world):
    a, oa = _replica(3)
    _step(a, oa, torch.randn(5, 8))
    before = {n: p.full_tensor().clone() for n, p in a.named_parameters()}
    load_fsdp_local_state(a, oa, fsdp_local_state(a, oa))  # every tensor already is the live storage: a no-op
    assert all(torch.equal(before[n], p.full_tensor()) for n, p in a.named_parameters())
    bad = fsdp_local_state(a, oa)
    bad["params"]["0.weight"] = torch.zeros(3, 3)
    with pytest.raises(ValueError, match="different sharding"):
        load_fsdp_local_state(a, oa, bad)
