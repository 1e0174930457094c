"""A small BPE vocabulary written in the layout of CLIP's merges file (shared by the tokenizer tests and the GPU text_probing test)."""

import gzip


def learn_merges(words, n):
    """Plain BPE training (most frequent adjacent pair first) — only to get a realistic, ordered merges list."""
    from collections import Counter

    vocab = Counter({tuple(w[:-1]) + (w[-1] + "</w>",): 1 for w in words})
    merges = []
    for _ in range(n):
        pairs = Counter()
        for w, c in vocab.items():
            for a, b in zip(w[:-1], w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda pr: pairs[pr])
        merges.append(best)
        nv = Counter()
        for w, c in vocab.items():
            out, i = [], 0
            while i < len(w):
                if i < len(w) - 1 and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            nv[tuple(out)] += c
        vocab = nv
    return merges


WORDS = ["photo", "of", "a", "dog", "cat", "the", "an", "image", "picture", "red", "blue", "car", "tree", "house", "hello",
         "world", "dogs", "it's", "2", "cats"]


def write_bpe_file(path, merges, trailing=5):
    """bpe_simple_vocab_16e6.txt.gz layout: a header line, one "left right" merge per line; the real file carries more
    lines than the 48 894 merges CLIP uses (the loader slices), mimicked here by lines past an explicit limit."""
    with gzip.open(path, "wt", encoding="utf-8") as f:
        f.write('"bpe_simple_vocab_16e6.txt#version: 0.2\n')
        for a, b in merges:
            f.write(f"{a} {b}\n")
        for i in range(trailing):
            f.write(f"zz{i} qq{i}\n")
    return path
