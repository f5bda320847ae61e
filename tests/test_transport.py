"""Lossless narrow transport of a collated batch (swarm_ode_b200.data.PackedBatch): the packed form must reproduce
``batch.to(device)`` of the reference (scripts/train_gde.py:475) bit for bit."""
import pytest
import torch

import swarm_ode_b200 as S
from swarm_ode_b200.data import PackedBatch, narrowest_exact_dtype, pack_bits, unpack_bits_host


def test_narrowest_exact_dtype_picks_by_round_trip():
    assert narrowest_exact_dtype(torch.tensor([[0.0, 1.0, 35.0, 255.0]]))[2] == "u8"
    assert narrowest_exact_dtype(torch.tensor([[0.0, 256.0]]))[2] == "i16"
    assert narrowest_exact_dtype(torch.tensor([[-1.0, 7.0]]))[2] == "i16"
    assert narrowest_exact_dtype(torch.tensor([[0.5, -0.25, 1024.0]]))[2] == "f16"
    assert narrowest_exact_dtype(torch.tensor([[0.1, 3.0]]))[2] == "f32"            # 0.1f is not a half
    assert narrowest_exact_dtype(torch.tensor([[1.0, float("nan")]]))[2] == "f32"
    assert narrowest_exact_dtype(torch.tensor([[1.0, float("inf")]]))[2] == "f32"
    assert narrowest_exact_dtype(torch.tensor([[-0.0, 3.0]]))[2] == "f16"           # the sign of zero is a bit too
    assert narrowest_exact_dtype(torch.tensor([[70000.0, 3.0]]))[2] == "f32"
    assert narrowest_exact_dtype(torch.empty(0, 4))[2] == "u8"


def test_warehouse_observations_pack_to_bytes():
    """The synthetic warehouse batch (flags and un-normalised grid coordinates) is integer valued."""
    batch, nxt = S.synthetic.warehouse_batch(16, seed=3)
    pb = PackedBatch(batch, nxt, bits=False)
    assert pb.kind_name == "u8" and pb.edges_int32 and pb.batch is None
    full = sum(t.numel() * t.element_size() for t in (batch.x, batch.edge_index, batch.batch, batch.ptr, batch.is_current_agent, nxt))
    assert pb.nbytes < 0.27 * full
    assert torch.equal(pb.x_packed.to(torch.float32), batch.x)
    assert torch.equal(pb.edge_index.to(torch.int64), batch.edge_index)
    # default: column-wise bit widths (flags 1 bit, grid coordinates 5 bits)
    pbb = PackedBatch(batch, nxt)
    assert pbb.kind_name == "bits" and pbb.x_packed.shape[1] % 16 == 0 and pbb.x_packed.shape[1] <= 128
    assert pbb.nbytes < 0.09 * full
    assert torch.equal(unpack_bits_host(pbb.x_packed, pbb.bit_offsets), batch.x)


def test_pack_bits_round_trip_and_refusals():
    g = torch.Generator().manual_seed(1)
    hi = torch.tensor([1, 2, 3, 4, 7, 8, 15, 16, 31, 100, 255, 0, 1, 1, 1, 200, 5], dtype=torch.float32)
    x = torch.floor(torch.rand(257, hi.numel(), generator=g) * (hi + 1.0)).clamp_max(hi)
    x[0] = hi                                                     # every column reaches its maximum
    packed, off = pack_bits(x)
    widths = (off[1:] - off[:-1]).tolist()
    assert widths == [1, 2, 2, 3, 3, 4, 4, 5, 5, 7, 8, 0, 1, 1, 1, 8, 3]
    assert packed.shape[1] % 16 == 0 and packed.shape[1] * 8 >= int(off[-1]) + 8
    assert torch.equal(unpack_bits_host(packed, off), x)
    for bad in (torch.tensor([[0.5, 1.0]]), torch.tensor([[-1.0, 1.0]]), torch.tensor([[256.0, 1.0]]), torch.tensor([[-0.0, 1.0]]),
                torch.tensor([[float("nan"), 1.0]]), torch.empty(0, 3)):
        assert pack_bits(bad) is None
    frac = S.Batch.from_data_list([S.Data(x=torch.tensor([[0.5, 1.0]]), edge_index=torch.empty((2, 0), dtype=torch.long))])
    assert PackedBatch(frac).kind_name == "f16"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bits", "u8", "i16", "f16", "f32"])
def test_packed_batch_arrives_bit_exact(cuda, kind):
    batch, nxt = S.synthetic.warehouse_batch(9, num_agvs=5, num_pickers=3, seed=11)
    if kind == "u8":
        batch.x[:, :] = torch.randint(128, 256, batch.x.shape).float()     # eight bits everywhere: bit packing cannot win
    if kind == "i16":
        batch.x[3, 2] = -7.0
    elif kind == "f16":
        batch.x[3, 2] = 0.375
        batch.x[5, 1] = -0.0
    elif kind == "f32":
        batch.x[3, 2] = 0.1
    pb = PackedBatch(batch, nxt).pin_memory()
    assert pb.kind_name == kind
    got, gn = pb.to(cuda, non_blocking=True)
    want = S.Batch(x=batch.x.clone(), edge_index=batch.edge_index.clone())
    want.batch, want.ptr, want.is_current_agent = batch.batch, batch.ptr, batch.is_current_agent
    torch.cuda.synchronize()
    assert got.x.dtype == torch.float32 and got.edge_index.dtype == torch.int64 and got.batch.dtype == torch.int64
    assert torch.equal(got.x.cpu().view(torch.int32), batch.x.view(torch.int32))          # bits, not values
    assert torch.equal(got.edge_index.cpu(), batch.edge_index)
    assert torch.equal(got.batch.cpu(), batch.batch)
    assert torch.equal(got.ptr.cpu(), batch.ptr)
    assert torch.equal(got.is_current_agent.cpu(), batch.is_current_agent)
    assert torch.equal(gn.cpu(), nxt)
    assert got.num_graphs == batch.num_graphs and got.max_graph_nodes == batch.max_graph_nodes


@pytest.mark.gpu
def test_packed_batch_ragged_sizes_and_empty_graphs(cuda):
    """Element counts that are not multiples of 16 (vector tail), graphs of different sizes, an empty graph in ptr."""
    g = torch.Generator().manual_seed(0)
    sizes = [3, 0, 17, 1, 40]
    graphs = []
    for n in sizes:
        x = torch.randint(0, 40, (n, 13), generator=g).float()
        ei = torch.randint(0, max(n, 1), (2, 2 * n), generator=g) if n > 1 else torch.empty((2, 0), dtype=torch.long)
        graphs.append(S.Data(x=x, edge_index=ei, is_current_agent=torch.ones(n, dtype=torch.bool)))
    batch = S.Batch.from_data_list(graphs)
    got = PackedBatch(batch).to(cuda)
    torch.cuda.synchronize()
    assert torch.equal(got.x.cpu(), batch.x) and torch.equal(got.edge_index.cpu(), batch.edge_index)
    assert torch.equal(got.batch.cpu(), batch.batch)
