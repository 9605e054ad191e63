"""Host logic of the N2 row (no GPU): the vectorised id join equals the reference's dict join on the golden fixture."""
import torch

from conftest import load_golden


def test_id_join_matches_the_reference_alignment(rs):
    a = load_golden("alignment.pt")
    src = rs.alignment.id_join(a["pretrained_ids"], a["item_ids"])
    hit = src >= 0
    assert int(hit.sum()) == 531
    torch.manual_seed(a["seed"])
    want = torch.randn(len(a["item_ids"]) + 1, a["dim"]) * 0.01
    want[0] = 0
    want[1:][hit] = a["pretrained"][src[hit]]
    assert torch.equal(want, a["aligned"])
    # a repeated exported id resolves to its LAST row (dict semantics of the reference)
    ids = ["b", "a", "c", "a"]
    assert rs.alignment.id_join(ids, ["a", "z", "c"]).tolist() == [3, -1, 2]
    assert rs.alignment.id_join(torch.tensor([5, 7]), ["7", "5", "05"]).tolist() == [1, 0, -1]
    assert rs.alignment.id_join([], ["a"]).tolist() == [-1]


def test_alignment_has_no_cpu_path(rs):
    import pytest
    with pytest.raises(RuntimeError, match="CUDA"):
        rs.alignment.align_pretrained(None, None, ["a"], 8, "cpu")
