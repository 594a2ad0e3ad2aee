from hypothesis import given, settings, strategies as st
import torch

from nanorlhf_b200.utils.batching import create_batches, strip_common_padding


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(1, 500), min_size=1, max_size=60), st.integers(500, 4000), st.sampled_from(["padded", "packed"]))
def test_create_batches_invariants(lengths, budget, mode):
    batches = create_batches(lengths, budget, mode)
    flat = [i for b in batches for i in b]
    assert sorted(flat) == list(range(len(lengths)))                       # a partition
    prev = 0
    for b in batches:
        ls = [lengths[i] for i in b]
        assert ls == sorted(ls) and ls[0] >= prev                           # ascending lengths
        prev = ls[-1]
        cost = max(ls) * len(ls) if mode == "padded" else sum(ls)
        assert cost <= budget or len(b) == 1                                 # budget respected


def test_strip_common_padding():
    pad = 0
    q = torch.tensor([[0, 0, 5, 6], [0, 7, 8, 9]])
    r = torch.tensor([[1, 2, 0, 0], [3, 0, 0, 0]])
    q2, r2 = strip_common_padding(q, r, pad)
    assert q2.tolist() == [[0, 5, 6], [7, 8, 9]] and r2.tolist() == [[1, 2], [3, 0]]
