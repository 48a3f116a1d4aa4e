"""CPU: what the reference's callers do to the decoder module outside the hot path: it must survive
``.share_memory()`` (src/DF_Prior.py:108-110), ``copy.deepcopy`` (src/Tracker.py:144), pickling across the spawned Mapper /
Tracker processes, ``parameters()`` grouping by submodule (src/Mapper.py:364-371) and a ``state_dict`` round trip with the
reference's 79 tensor names (src/utils/Logger.py:26, DF_Prior.load_pretrain :205-218).  No kernel is called here."""
import copy
import pickle

import torch

import attentive_dfprior_b200 as A
from oracle import dfprior_oracle as O


def test_state_dict_has_the_reference_names_and_sizes():
    model = A.DF()
    sd = model.state_dict()
    ref = O.init_decoder_params(0)                       # the oracle's parameter dict uses the reference's names
    assert list(sd.keys()) == list(ref.keys()) and len(sd) == 79
    assert all(tuple(sd[k].shape) == tuple(ref[k].shape) for k in sd)
    assert sum(v.numel() for v in sd.values()) == 86029


def test_share_memory_deepcopy_pickle_and_grouping():
    model = A.DF()
    model.bound = torch.tensor([[0.0, 1.0], [0.0, 2.0], [0.0, 3.0]], dtype=torch.float64)
    model.share_memory()
    assert all(p.is_shared() for p in model.parameters())
    m2 = copy.deepcopy(model)
    m3 = pickle.loads(pickle.dumps(model))
    for other in (m2, m3):
        assert torch.equal(other.bound, model.bound)
        for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
            assert torch.equal(a, b), k
        assert other.low_decoder.fc_c[0].weight.data_ptr() != model.low_decoder.fc_c[0].weight.data_ptr()
    groups = {name: list(getattr(model, name).parameters()) for name in ("low_decoder", "high_decoder", "color_decoder", "mlp")}
    assert [len(groups[n]) for n in ("low_decoder", "high_decoder", "color_decoder", "mlp")] == [23, 23, 23, 10]
    assert len(list(model.parameters())) == 79
    m4 = A.DF()
    m4.load_state_dict(model.state_dict())
    assert all(torch.equal(a, b) for a, b in zip(model.state_dict().values(), m4.state_dict().values()))
    for p in m4.high_decoder.parameters():
        p.requires_grad_(False)                           # a frozen decoder stays an ordinary module
    assert sum(p.requires_grad for p in m4.parameters()) == 79 - 23


def test_frozen_flags_follow_requires_grad():
    """DFP_FLAG_FROZEN_* are derived from which of the 79 tensors need a gradient (state_dict order: low, high, colour, mlp)."""
    from attentive_dfprior_b200 import _lib as L
    from attentive_dfprior_b200.functional import _frozen_flags
    all_on = [True] * 79
    assert _frozen_flags(all_on, True) == 0
    assert _frozen_flags(all_on, False) == 0                                  # no backward at all: the flags mean nothing
    need = list(all_on)
    need[0:23] = [False] * 23
    assert _frozen_flags(need, True) == L.FLAG_FROZEN[0]
    need[23:46] = [False] * 23
    assert _frozen_flags(need, True) == L.FLAG_FROZEN[0] | L.FLAG_FROZEN[1]
    need[45] = True                                                           # one tensor of the high decoder still trained
    assert _frozen_flags(need, True) == L.FLAG_FROZEN[0]
    assert _frozen_flags([False] * 79, True) == L.FLAG_FROZEN[0] | L.FLAG_FROZEN[1] | L.FLAG_FROZEN[2]     # Tracker: only d(rays)
    assert (L.FLAG_SAVE_ACT | L.FLAG_SIMT) & (L.FLAG_FROZEN[0] | L.FLAG_FROZEN[1] | L.FLAG_FROZEN[2]) == 0
