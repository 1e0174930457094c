"""The torch-CPU port of the reference path (oracle/ref_port.py, the thing bench.py times as the CPU baseline)
against the fixtures recorded from the imported reference (oracle/make_golden.py)."""

import warnings

import numpy as np
import pytest
import torch

from oracle import ref_port as rp


def bits_of(t):
    return t.view(torch.int16).numpy().view(np.uint16)


class _Emit(torch.nn.Module):
    def forward(self, x):
        return x


NAMES = sorted(rp.AGGREGATORS) + ["aggregate_transformer_special_token"]


@pytest.mark.parametrize("name", NAMES)
def test_port_hook_sweep_is_bit_identical_to_reference(golden, name):
    z = np.load(golden / f"collect_{name}.npz")
    fn = rp.AGGREGATORS.get(name) or rp.get_aggregate_transformer_special_token(0)
    model = torch.nn.Sequential()
    model.add_module("probe", _Emit())
    maps = [torch.from_numpy(z[f"map{i}"]) for i in range(int(z["n_batches"]))]
    for i, m in enumerate(maps):
        assert torch.equal(fn(m), torch.from_numpy(z[f"agg{i}"]))
    st = rp.sweep(model, [(m, None) for m in maps], ["probe"], fn, int(z["k"]))["probe"]
    assert (bits_of(st.activations) == z["ref_bits"]).all()
    assert (st.sample_ids.numpy() == z["ref_ids"]).all()  # same torch.topk => same tie order


def test_port_reference_kat(golden):
    z = np.load(golden / "actmax_kat.npz")
    am = rp.ActMaxPort(5, 3)
    am.update(torch.from_numpy(z["acts1"]), torch.tensor([0, 1]))
    am.update(torch.from_numpy(z["acts2"]), torch.tensor([2, 3]))
    assert am.sample_ids[0].tolist() == [2, 3, 1, 0, -1]
    assert (bits_of(am.activations) == z["ref_bits"]).all() and (am.sample_ids.numpy() == z["ref_ids"]).all()


def test_port_scores(golden):
    z = np.load(golden / "scores.npz")
    t = lambda k: torch.from_numpy(z[k])  # noqa: E731
    assert torch.equal(rp.clarity_score(t("V")), t("clarity"))
    assert torch.equal(rp.similarity_score(t("sim_x"), t("sim_y")), t("sim_xy"))
    assert torch.equal(rp.similarity_score(t("sim_x"), t("sim_y2")), t("sim_xy2"))  # C == D: no transpose
    assert torch.equal(rp.similarity_score(t("sim_x3"), t("sim_y")), t("sim_x3y"))  # equal shapes: row-wise
    assert torch.equal(rp.redundancy_score(t("red_in")), t("red"))
    assert torch.equal(rp.redundancy_score(t("red2_in")), t("red2"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        poly = rp.polysemanticity_score(t("P"))
        poly_nr = rp.polysemanticity_score(t("P"), replace_empty_clusters=False)
    assert poly.dtype == torch.float64
    np.testing.assert_allclose(poly.numpy(), z["poly"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(poly_nr.numpy(), z["poly_noreplace"], rtol=1e-12, atol=1e-12)


def test_port_concept_db_negative_index_aliases_last_image():
    st = rp.ActMaxPort(3, 2)
    st.update(torch.tensor([[1.0, -1.0]]), torch.tensor([0]))
    emb = torch.arange(8.0).reshape(4, 2)
    db = rp.concept_db({"l": st}, emb)["l"]
    assert db.shape == (2, 3, 2)
    assert torch.equal(db[1], emb[[-1, -1, -1]]) and torch.equal(db[0, 0], emb[0])
