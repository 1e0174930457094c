"""The BPE tokenizer (open_clip SimpleTokenizer scheme) on a synthetic merges list: CLIP's real merges file is not
available offline, so the algorithm is pinned on a vocabulary small enough to verify by hand."""

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
