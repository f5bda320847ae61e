"""swarm_ode_b200.hetero.MultiAgentGraphConverter against golden vectors produced by the REFERENCE's own class
(scripts/run_gnode.py:1040-1326, executed from its source text by scripts/make_golden.py): node features and the six
relations bit for bit, the reference's failure modes as exceptions of the same type, and the stale state of a reused
converter.  Host-side integer work: no GPU needed."""
import os

import numpy as np
import pytest
import torch

from swarm_ode_b200.hetero import EDGE_TYPES, HeteroData, MultiAgentGraphConverter

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multi_agent_converter.npz"))
SHAPES = {"small": (4, 3), "medium": (19, 9)}
EXC = {"KeyError": KeyError, "ValueError": ValueError, "TypeError": TypeError}


def _racks(tag):
    return [tuple(int(v) for v in r) for r in GOLD[f"{tag}/racks"]]


def _check(tag, conv, racks=None):
    obs = GOLD[f"{tag}/obs"]
    racks = _racks(tag) if racks is None else racks
    if f"{tag}/raises" in GOLD.files:
        with pytest.raises(EXC[str(GOLD[f'{tag}/raises'])]):
            conv._build_graph_from_observation(obs, racks)
        return
    d = conv._build_graph_from_observation(obs, racks)
    for k in ("agv", "picker", "location"):
        want = GOLD[f"{tag}/x/{k}"]
        got = d[k].x.numpy()
        assert got.dtype == want.dtype and got.shape == want.shape and np.array_equal(got, want), (tag, k)
    for et in EDGE_TYPES:
        want = GOLD[f"{tag}/edge/" + "__".join(et)]
        got = d[et].edge_index
        assert got.dtype == torch.int64 and tuple(got.shape) == want.shape, (tag, et, tuple(got.shape), want.shape)
        assert np.array_equal(got.numpy(), want), (tag, et)


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("case", ["idle", "one_agv_target", "agv_target_off_rack", "nothing_requested", "two_agv_targets",
                                  "picker_target"])
def test_single_call_matches_reference(shape, case):
    _check(f"{shape}/{case}", MultiAgentGraphConverter(*SHAPES[shape]))


@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_reused_converter_keeps_the_edges_of_its_first_observation(shape):
    conv = MultiAgentGraphConverter(*SHAPES[shape])
    _check(f"{shape}/reuse_first", conv)
    _check(f"{shape}/reuse_second", conv)
    # fresh=True: the second observation's own edges = a new reference converter on that observation (golden 'idle'-like)
    fresh = MultiAgentGraphConverter(*SHAPES[shape], fresh=True)
    fresh._build_graph_from_observation(GOLD[f"{shape}/reuse_first/obs"], _racks(f"{shape}/reuse_first"))
    second = fresh._build_graph_from_observation(GOLD[f"{shape}/reuse_second/obs"], _racks(f"{shape}/reuse_second"))
    alone = MultiAgentGraphConverter(*SHAPES[shape])._build_graph_from_observation(GOLD[f"{shape}/reuse_second/obs"],
                                                                                   _racks(f"{shape}/reuse_second"))
    for et in EDGE_TYPES:
        assert torch.equal(second[et].edge_index, alone[et].edge_index)
    stale = GOLD[f"{shape}/reuse_second/edge/agv__targets__location"]
    assert not np.array_equal(stale, alone["agv", "targets", "location"].edge_index.numpy())   # the quirk is real


def test_ndarray_rack_rows_raise_like_the_reference():
    tag = "small/ndarray_racks"
    _check(tag, MultiAgentGraphConverter(*SHAPES["small"]), racks=np.asarray(GOLD[f"{tag}/racks"]))


def test_hetero_batch_is_a_disjoint_union():
    """HeteroData.from_data_list: features concatenated, every relation shifted by the per-type node offsets."""
    convs = [MultiAgentGraphConverter(*SHAPES["small"]) for _ in range(3)]
    tags = ["small/idle", "small/one_agv_target", "small/nothing_requested"]
    ds = [c._build_graph_from_observation(GOLD[f"{t}/obs"], _racks(t)) for c, t in zip(convs, tags)]
    b = HeteroData.from_data_list(ds)
    assert b.num_graphs == 3
    for nt in ("agv", "picker", "location"):
        assert torch.equal(b[nt].x, torch.cat([d[nt].x for d in ds]))
        assert b[nt].ptr.tolist() == np.cumsum([0] + [d[nt].x.shape[0] for d in ds]).tolist()
    for et in EDGE_TYPES:
        src, _r, dst = et
        parts = [d[et].edge_index + torch.tensor([[int(b[src].ptr[g])], [int(b[dst].ptr[g])]]) for g, d in enumerate(ds)]
        assert torch.equal(b[et].edge_index, torch.cat(parts, dim=1))
