"""The BPE tokenizer (open_clip SimpleTokenizer scheme) on a synthetic merges list: CLIP's real merges file is not
available offline, so the algorithm is pinned on a vocabulary small enough to verify by hand, and the loader is pinned on a
file written in the real file's format (gzip, header line, one merge per line, unused trailing lines) against HF
transformers' CLIPTokenizer reading the equivalent vocab.json / merges.txt. The SigLIP sentencepiece wrapper is exercised
on a toy model trained in the test."""

import gzip
import json

import pytest
import torch

from semanticlens_b200.foundation_models import text as T

MERGES = [("h", "e"), ("l", "l"), ("he", "ll"), ("hell", "o</w>"), ("w", "o"), ("r", "l"), ("wo", "rl"), ("worl", "d</w>")]


def test_vocabulary_layout_matches_clip_scheme():
    tok = T.SimpleTokenizer(merges=MERGES, context_length=8)
    # 256 byte symbols, 256 word-final byte symbols, the merges, then <start_of_text>, <end_of_text>
    assert tok.vocab_size == 512 + len(MERGES) + 2
    assert (tok.sot, tok.eot) == (tok.vocab_size - 2, tok.vocab_size - 1)
    assert tok.encoder["!"] == 0 and tok.encoder["!</w>"] == 256  # "!" is the first printable byte in the byte table
    assert tok.encoder["hello</w>"] == 512 + 3 and tok.encoder["world</w>"] == 512 + 7


def test_encoding_lowercases_cleans_whitespace_and_merges():
    tok = T.SimpleTokenizer(merges=MERGES, context_length=8)
    ids = tok(["Hello   WORLD!", "", "he"])
    assert ids.dtype == torch.int64 and ids.shape == (3, 8)
    assert ids[0].tolist() == [tok.sot, tok.encoder["hello</w>"], tok.encoder["world</w>"], tok.encoder["!</w>"], tok.eot, 0, 0, 0]
    assert ids[1].tolist() == [tok.sot, tok.eot, 0, 0, 0, 0, 0, 0]
    # "he" alone: the merge ("h", "e") does not apply to ("h", "e</w>")
    assert ids[2].tolist()[:4] == [tok.sot, tok.encoder["h"], tok.encoder["e</w>"], tok.eot]


def test_truncation_keeps_end_of_text_as_the_argmax():
    tok = T.SimpleTokenizer(merges=MERGES, context_length=6)
    ids = tok("hello " * 20)
    assert ids.shape == (1, 6) and ids[0, 0] == tok.sot and ids[0, -1] == tok.eot
    assert int(ids[0].argmax()) == 5  # EOT pooling position


def test_missing_merges_file_is_reported():
    with pytest.raises(FileNotFoundError, match="bpe_simple_vocab"):
        T.SimpleTokenizer(bpe_path="/nonexistent/bpe.txt.gz")


from tests.bpe_fixture import WORDS, learn_merges, write_bpe_file  # noqa: E402


def test_merges_file_loader_against_hf_clip_tokenizer(tmp_path, monkeypatch):
    transformers = pytest.importorskip("transformers")
    merges = learn_merges(WORDS, 60)
    gz = write_bpe_file(tmp_path / "bpe_simple_vocab_16e6.txt.gz", merges, trailing=0)
    tok = T.SimpleTokenizer(bpe_path=str(gz), context_length=16)
    assert tok.vocab_size == 512 + len(merges) + 2
    assert list(tok.bpe_ranks)[:3] == merges[:3]
    # the same vocabulary for HF's CLIPTokenizer (its special tokens are spelled differently)
    vocab = dict(tok.encoder)
    vocab["<|startoftext|>"] = vocab.pop("<start_of_text>")
    vocab["<|endoftext|>"] = vocab.pop("<end_of_text>")
    (tmp_path / "vocab.json").write_text(json.dumps(vocab))
    (tmp_path / "merges.txt").write_text("#version: 0.2\n" + "".join(f"{a} {b}\n" for a, b in merges))
    hf = transformers.CLIPTokenizer(str(tmp_path / "vocab.json"), str(tmp_path / "merges.txt"))
    texts = ["a photo of a dog", "An IMAGE   of the red car!", "it's 2 cats, hello-world", "", "tree house picture of blue dogs"]
    for text in texts:
        want = hf(text)["input_ids"]
        got = tok(text)[0]
        n = len(want)
        assert got[:n].tolist() == want and (got[n:] == 0).all(), text
    # SLB_CLIP_BPE is honoured when no path is passed
    monkeypatch.setenv("SLB_CLIP_BPE", str(gz))
    assert T.SimpleTokenizer().vocab_size == tok.vocab_size


def test_merges_file_loader_ignores_lines_past_clip_vocabulary(tmp_path, monkeypatch):
    """CLIP uses lines 1 .. 48 894 of the file; anything after is ignored. Checked with a lowered limit."""
    merges = learn_merges(WORDS, 40)
    gz = write_bpe_file(tmp_path / "bpe.txt.gz", merges, trailing=7)
    monkeypatch.setattr(T, "CLIP_MERGES", len(merges))
    tok = T.SimpleTokenizer(bpe_path=str(gz))
    assert tok.vocab_size == 512 + len(merges) + 2 and ("zz0", "qq0") not in tok.bpe_ranks
    monkeypatch.setattr(T, "CLIP_MERGES", 49152 - 256 - 2)
    assert ("zz0", "qq0") in T.SimpleTokenizer(bpe_path=str(gz)).bpe_ranks


def test_sentencepiece_tokenizer_mechanics(tmp_path):
    spm = pytest.importorskip("sentencepiece")
    corpus = tmp_path / "corpus.txt"
    corpus.write_text("\n".join(" ".join(WORDS[i % len(WORDS)] for i in range(j, j + 6)) for j in range(200)))
    spm.SentencePieceTrainer.train(input=str(corpus), model_prefix=str(tmp_path / "toy"), vocab_size=40, model_type="unigram",
                                   pad_id=0, unk_id=2, bos_id=-1, eos_id=1, minloglevel=2)
    tok = T.SentencePieceTokenizer(str(tmp_path / "toy.model"), context_length=12)
    assert T.SentencePieceTokenizer.canonicalize("  A PHOTO, of a   dog!! ") == "a photo of a dog"
    ids = tok(["A photo, of a dog!", "", "dog " * 40])
    assert ids.shape == (3, 12) and ids.dtype == torch.int64
    body = tok.sp.encode("a photo of a dog")
    assert ids[0, : len(body)].tolist() == body and ids[0, len(body)] == tok.eos and (ids[0, len(body) + 1 :] == 1).all()
    assert ids[1, 0] == tok.eos and (ids[1, 1:] == 1).all()          # empty prompt: </s> then padding (= </s>)
    assert (ids[2] != 1).all()                                        # truncated to the context length
    with pytest.raises(FileNotFoundError, match="sentencepiece"):
        T.SentencePieceTokenizer("/nonexistent/spiece.model")
