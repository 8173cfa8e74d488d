"""Generates tests/golden/tokenizer_golden.json: token ids produced by the REFERENCE's own tokenizer
(/root/reference/tinyfusers/tokenizer/clip.py:10-75, imported unmodified) for a set of prompts. Test infrastructure only.

The reference downloads OpenAI's bpe_simple_vocab_16e6.txt.gz at import time (clip.py:7-8); there is no network here and the
file is not on this machine, so the golden is taken on a merges file this script LEARNS (plain byte-pair merging over a fixed
English corpus, 600 rules, written in the same format: a header line, then one "a b" rule per line). Both tokenizers read the
same file, so every code path of `encode` is pinned: lower-casing and whitespace cleaning, the pre-token regular expression
(contractions, punctuation glued to words), byte -> unicode mapping of non-ASCII input, rank-ordered merging with repeated
symbols, truncation to 75 ids, BOS / EOS framing and padding. The merges file itself is stored in the golden.

    python oracle/make_tokenizer_golden.py
"""
import collections
import gzip
import json
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

CORPUS = """
a photograph of an astronaut riding a horse on mars, highly detailed, sharp focus, dramatic lighting. the quick brown fox
jumps over the lazy dog. a horse sized cat eating a bagel. an oil painting of a lighthouse at sunset, trending on artstation.
it's a beautiful day and i'm sure you'll love what we've done; they're here, she'd say, but the cat's toy wasn't.
stable diffusion renders images from text prompts with a latent denoising model conditioned on text embeddings.
the the the and and of of in in to to a a is is that that it it for for was was on on with with as as be be
painting paintings painted painter detailed detail details lighting light lights rendering render renders rendered
mountain mountains river rivers forest forests city cities street streets portrait portraits landscape landscapes
""" * 3

PROMPTS = [
    "a horse sized cat eating a bagel",
    "",
    "A Photograph of an Astronaut   riding a horse on Mars,\thighly detailed!!",
    "it's what we've done; they're here, she'd say: i'm sure you'll",
    "café naïve über — résumé 日本語 \U0001f600",
    "<|startoftext|> the cat <|endoftext|> sat",
    "aaaa aaaaa bababab mississippi lighthouselighthouse",
    "1234567890 3.14159 #hashtag @user http://example.com/a?b=c&d=e",
    " ".join(["painting of a forest river in the mountains"] * 12),        # > 75 ids: truncation
    "THE QUICK BROWN FOX JUMPS OVER THE LAZY DOG.",
]


def bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def learn_merges(text, n_rules):
    """Plain BPE training over whitespace-separated words in the byte -> unicode alphabet, end-of-word marker '</w>'."""
    b2u = bytes_to_unicode()
    words = collections.Counter()
    for w in text.lower().split():
        sym = [b2u[b] for b in w.encode("utf-8")]
        sym[-1] += "</w>"
        words[tuple(sym)] += 1
    rules = []
    for _ in range(n_rules):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w[:-1], w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        best = max(sorted(pairs), key=lambda p: pairs[p])      # deterministic tie-break
        rules.append(best)
        merged = {}
        for w, c in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and (w[i], w[i + 1]) == best:
                    out.append(w[i] + w[i + 1])
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            merged[tuple(out)] = merged.get(tuple(out), 0) + c
        words = collections.Counter(merged)
    return rules


def main():
    rules = learn_merges(CORPUS, 600)
    merges_text = "#version: learned by oracle/make_tokenizer_golden.py\n" + "\n".join(f"{a} {b}" for a, b in rules) + "\n"
    tmp = tempfile.mkdtemp()
    path = os.path.join(tmp, "learned_bpe.txt.gz")
    with gzip.open(path, "wb") as fh:
        fh.write(merges_text.encode("utf-8"))
    # the reference module imports tinygrad.helpers.fetch and calls it at class-definition time (default argument)
    tg = types.ModuleType("tinygrad")
    tgh = types.ModuleType("tinygrad.helpers")
    tgh.fetch = lambda url, name=None: path
    tg.helpers = tgh
    sys.modules["tinygrad"], sys.modules["tinygrad.helpers"] = tg, tgh
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_clip_tokenizer", os.path.join(REF, "tinyfusers", "tokenizer", "clip.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    tok = mod.ClipTokenizer(path)
    out = {"merges": merges_text, "cases": []}
    for ptxt in PROMPTS:
        try:
            ids = tok.encode(ptxt)
            out["cases"].append({"prompt": ptxt, "ids": [int(i) for i in ids]})
        except Exception as exc:      # recorded, not hidden: a prompt the reference cannot encode with this vocabulary
            out["cases"].append({"prompt": ptxt, "error": type(exc).__name__})
    with open(os.path.join(ROOT, "tests", "golden", "tokenizer_golden.json"), "w") as fh:
        json.dump(out, fh, ensure_ascii=True, indent=0)
    for c in out["cases"]:
        print(repr(c["prompt"][:50]), c.get("error") or (c["ids"][:12], sum(1 for i in c["ids"] if i != 49407)))


if __name__ == "__main__":
    main()
