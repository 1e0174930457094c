"""Host logic of `Lens` and the act-max cache on the CPU, in the style of the reference's own tests
(tests/test_lens.py:35-97, tests/component_visualization/test_activation_caching.py:32-91): mocks for the foundation
model and the visualizer, no kernel is launched."""

from pathlib import Path
from unittest import mock

import numpy as np
import pytest
import torch

from semanticlens_b200 import lens as L
from semanticlens_b200.component_visualization import aggregators as A
from semanticlens_b200.component_visualization.activation_caching import ActMax, ActMaxCache
from tests.collect_cases import TEXT_PROBE, FakeTextFM


@pytest.fixture
def fm():
    m = mock.MagicMock()
    m.device = "cpu"
    m.name = "fake-fm"
    m.encode_text.return_value = torch.randn(1, 128)
    return m


@pytest.fixture
def cv(tmp_path):
    m = mock.MagicMock()
    m.caching = True
    m.storage_dir = tmp_path
    m.device = "cpu"
    m.metadata = {"aggregation_fn_name": "agg", "n_collect": "3", "layer_names": "['layer1']", "dataset": "d", "model": "m"}
    m._compute_concept_db.return_value = {"layer1": torch.randn(4, 3, 8)}
    return m


def test_lens_moves_fm_and_names_it(fm):
    lens = L.Lens(fm=fm, device="cpu")
    assert lens.fm is fm
    fm.to.assert_called_with("cpu")
    bare = mock.MagicMock(spec=["device", "to"])
    bare.device = "cpu"
    L.Lens(bare)
    assert bare.name.startswith("MagicMock-")  # fallback name = ClassName-<hash>, names the cache directory


def test_concept_db_cache_miss_computes_and_saves(fm, cv):
    with mock.patch.object(L, "save_file") as save:
        db = L.Lens(fm).compute_concept_db(cv)
    cv._compute_concept_db.assert_called_once_with(fm)
    save.assert_called_once()
    assert "layer1" in db
    # file grammar of reference lens.py:308-316
    target = Path(save.call_args.kwargs["filename"])
    assert target == cv.storage_dir / "concept_database" / "fake-fm" / "concept_db-agg-3-['layer1'].safetensors"


def test_concept_db_cache_hit_loads(fm, cv):
    with mock.patch("pathlib.Path.exists", return_value=True), \
            mock.patch.object(L, "load_file", return_value={"layer1": "data_from_cache"}) as load:
        db = L.Lens(fm).compute_concept_db(cv)
    load.assert_called_once()
    cv._compute_concept_db.assert_not_called()
    assert db["layer1"] == "data_from_cache"


def test_concept_db_real_roundtrip(fm, cv):
    lens = L.Lens(fm)
    first = lens.compute_concept_db(cv)
    again = lens.compute_concept_db(cv)
    assert cv._compute_concept_db.call_count == 1
    assert torch.equal(first["layer1"], again["layer1"])


def test_no_caching_never_touches_disk(fm, cv):
    cv.caching = False
    with mock.patch.object(L, "save_file") as save, mock.patch.object(L, "load_file") as load:
        L.Lens(fm).compute_concept_db(cv, batch_size=7)
    cv._compute_concept_db.assert_called_once_with(fm, batch_size=7)
    save.assert_not_called()
    load.assert_not_called()


@pytest.mark.parametrize("case", sorted(TEXT_PROBE))
def test_embed_text_probes_matches_reference(golden, case):
    """Recorded from the reference's lens._embed_text_probes (oracle/make_golden.py:gen_text_probe), including its
    template-major list read back as (query, template) blocks."""
    queries, templates, bs = TEXT_PROBE[case]
    want = np.load(golden / "text_probe.npz")[case]
    got = L._embed_text_probes(FakeTextFM(), list(queries), list(templates) if templates else None, bs)
    assert got.shape == (len(queries), 16)
    np.testing.assert_array_equal(got.numpy(), want)


def test_template_regrouping_quirk_is_kept():
    """2 queries x 3 templates: prompt p = t * n_q + q is read as row p = q' * n_t + t' -> query 0 averages prompts 0..2."""
    fm = FakeTextFM()
    queries, templates = ["a dog", "sky"], ["a photo of {}", "{} texture", "close-up of {}"]
    got = L._embed_text_probes(fm, queries, templates, None)
    prompts = [t.format(q) for t in templates for q in queries]
    rows = fm.encode_text(fm.tokenize(prompts))
    base = fm.encode_text(fm.tokenize([t.format("") for t in templates]))
    want0 = (rows[0:3] - base).mean(0)
    assert torch.allclose(got[0], want0)


def test_probe_dispatch_tensor_and_dict():
    q = torch.randn(2, 8)
    with mock.patch.object(L, "similarity_score", side_effect=lambda a, b: a @ b.T) as sim:
        one = L._probe(q, torch.randn(5, 8))
        many = L._probe(q, {"a": torch.randn(5, 8), "b": torch.randn(3, 8)})
    assert one.shape == (2, 5) and many["a"].shape == (2, 5) and many["b"].shape == (2, 3)
    assert sim.call_count == 3


# ---- act-max cache files ----------------------------------------------------------------------------------------------
def _filled(n_latents=4, k=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    am = ActMax(n_collect=k, n_latents=n_latents)
    am.activations = torch.randn(n_latents, k, generator=g).bfloat16()
    am.sample_ids = torch.randint(0, 100, (n_latents, k), generator=g)
    return am


def test_actmax_store_load_roundtrip(tmp_path):
    am = _filled()
    f = tmp_path / "a.safetensors"
    am.store(f, metadata={"n_collect": "3", "n_latents": "4"})
    assert not list(tmp_path.glob("*.partial"))
    back = ActMax.load(f)
    assert back.n_collect == 3 and back.n_latents == 4 and back.is_setup
    assert torch.equal(back.activations.view(torch.int16), am.activations.view(torch.int16))
    assert torch.equal(back.sample_ids, am.sample_ids)
    am.store(tmp_path / "nometa.safetensors")
    with pytest.raises(ValueError, match="metadata"):
        ActMax.load(tmp_path / "nometa.safetensors")
    empty = ActMax(n_collect=3)
    empty.store(tmp_path / "never.safetensors")
    assert not (tmp_path / "never.safetensors").exists()


def test_actmaxcache_files_metadata_and_misses(tmp_path, golden):
    cache = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 3)
    cache.cache["0"], cache.cache["2"] = _filled(4, 3, 1), _filled(6, 3, 2)
    cache.store(tmp_path / "c")
    ours = sorted(p.name for p in (tmp_path / "c").iterdir())
    theirs = sorted(p.name for p in (golden / "cache_format").iterdir())
    assert ours == theirs == ["aggregate_conv_mean-3-0.safetensors", "aggregate_conv_mean-3-2.safetensors"]
    import safetensors

    for name in ours:
        with safetensors.safe_open(str(tmp_path / "c" / name), framework="pt") as a, \
                safetensors.safe_open(str(golden / "cache_format" / name), framework="pt") as b:
            assert sorted(a.keys()) == sorted(b.keys()) == ["activations", "sample_ids"]
            assert set(a.metadata()) == set(b.metadata())
            assert a.get_tensor("activations").dtype == b.get_tensor("activations").dtype == torch.bfloat16
            assert a.get_tensor("sample_ids").dtype == b.get_tensor("sample_ids").dtype == torch.int64
    assert list(cache.metadata) == ["aggregation_fn_name", "n_collect", "layer_names"]

    # a directory written by the REAL reference loads
    ref = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 3)
    ref.load(golden / "cache_format")
    assert ref.cache["0"].activations.shape == (4, 3) and ref.cache["2"].sample_ids.shape == (6, 3)

    # misses: directory, one layer's file, other n_collect, other aggregation — all FileNotFoundError, state untouched
    fresh = ActMaxCache(["0", "2"], A.aggregate_conv_mean, 3)
    with pytest.raises(FileNotFoundError):
        fresh.load(tmp_path / "absent")
    (tmp_path / "c" / "aggregate_conv_mean-3-2.safetensors").rename(tmp_path / "keep.safetensors")
    with pytest.raises(FileNotFoundError):
        fresh.load(tmp_path / "c")
    assert not fresh.cache["0"].is_setup  # all-or-nothing: layer "0" was not half-loaded
    (tmp_path / "keep.safetensors").rename(tmp_path / "c" / "aggregate_conv_mean-3-2.safetensors")
    with pytest.raises(FileNotFoundError):
        ActMaxCache(["0", "2"], A.aggregate_conv_mean, 4).load(tmp_path / "c")
    with pytest.raises(FileNotFoundError):
        ActMaxCache(["0", "2"], A.aggregate_conv_max, 3).load(tmp_path / "c")
    # a file whose NAME matches but whose header was written for another aggregation function
    src = tmp_path / "c" / "aggregate_conv_mean-3-0.safetensors"
    (tmp_path / "d").mkdir()
    for layer in ("0", "2"):
        (tmp_path / "d" / f"aggregate_conv_max-3-{layer}.safetensors").write_bytes(src.read_bytes())
    with pytest.raises(FileNotFoundError, match="does not match"):
        ActMaxCache(["0", "2"], A.aggregate_conv_max, 3).load(tmp_path / "d")
    with pytest.raises(ValueError, match="lambda"):
        ActMaxCache(["0"], lambda t: t, 3)


def test_openclip_refuses_silent_random_weights():
    from semanticlens_b200.foundation_models import OpenClip, SigLipV2

    with pytest.raises(ValueError, match="cannot download"):
        OpenClip("ViT-B-32", device="cpu", pretrained="laion2b_s34b_b79k")
    with pytest.raises(ValueError, match="cannot download"):
        SigLipV2(device="cpu")
    fm = SigLipV2(device="cpu", load_weights=False)
    assert fm.resize_mode == "squash" and OpenClip("ViT-B-32", device="cpu", load_weights=False).resize_mode == "shortest"


def test_vit_tower_rejects_a_checkpoint_of_another_architecture():
    from semanticlens_b200.foundation_models import vit

    sd = vit.random_state_dict(vit.CONFIGS["ViT-B-32"], 0)
    with pytest.raises(ValueError, match="wrong shape"):
        vit.VitTower(vit.CONFIGS["ViT-B-16"], sd, "cpu")


def test_squash_and_shortest_preprocess_geometry():
    """The host path for non-RGB images (PIL resize) against torchvision's transforms: Resize((S, S)) for SigLIP,
    Resize(S) + CenterCrop(S) for CLIP."""
    from PIL import Image
    from torchvision import transforms as T

    from semanticlens_b200.foundation_models.clip import _pil_to_chw_u8

    rng = np.random.default_rng(0)
    im = Image.fromarray(rng.integers(0, 256, (90, 131, 3), dtype=np.uint8))
    bic = T.InterpolationMode.BICUBIC
    squash = np.asarray(T.Resize((64, 64), interpolation=bic)(im)).transpose(2, 0, 1)
    crop = np.asarray(T.CenterCrop(64)(T.Resize(64, interpolation=bic)(im))).transpose(2, 0, 1)
    assert np.array_equal(_pil_to_chw_u8(im, 64, "squash"), squash)
    assert np.array_equal(_pil_to_chw_u8(im, 64, "shortest"), crop)
    assert not np.array_equal(squash, crop)
