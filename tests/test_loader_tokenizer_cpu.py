"""Host-side glue next to the hot path: the torch-zip checkpoint reader and the CLIP BPE
tokenizer, with the reference's interfaces (storage/unpicker.py:74 load_weights, tokenizer/clip.py:10-75)."""
import collections
import gzip
import os

import numpy as np
import pytest
import torch


def test_load_weights_roundtrip(tmp_path):
    from tinyfusers_b200.storage.unpicker import load_weights
    g = torch.Generator().manual_seed(0)
    base = torch.randn(6, 8, generator=g)
    sd = collections.OrderedDict()
    sd["a.weight"] = torch.randn(4, 3, 3, 3, generator=g)
    sd["b.half"] = torch.randn(5, 7, generator=g).half()                  # the reference decodes these as fp32 words
    sd["c.view"] = base[1:5, 2:6]                                         # storage offset + non-contiguous strides
    sd["c.t"] = base.t()                                                  # transposed view of the same storage
    sd["d.scalar"] = torch.tensor(3.5)
    sd["e.long"] = torch.arange(10)
    sd["f.bf16"] = torch.randn(3, 4, generator=g).bfloat16()
    sd["g.param"] = torch.nn.Parameter(torch.randn(2, 2, generator=g))
    path = os.path.join(tmp_path, "ckpt.pt")
    torch.save({"state_dict": sd, "global_step": 7}, path)
    out = load_weights(path)
    assert out["global_step"] == 7
    got = out["state_dict"]
    assert list(got.keys()) == list(sd.keys())
    for k, v in sd.items():
        ref = v.detach()
        if ref.is_floating_point():
            assert got[k].dtype == np.float32, k
            np.testing.assert_array_equal(got[k], ref.float().numpy(), err_msg=k)
        else:
            np.testing.assert_array_equal(got[k], ref.numpy(), err_msg=k)
        assert got[k].flags["C_CONTIGUOUS"]


def test_load_weights_feeds_update_state(tmp_path):
    """The arrays go straight into the model tree (reference flow: example/sd1.py:40-41)."""
    from tinyfusers_b200.ff.linear import Linear
    from tinyfusers_b200.storage.state import update_state
    from tinyfusers_b200.storage.unpicker import load_weights
    w, b = torch.randn(16, 8), torch.randn(16)
    path = os.path.join(tmp_path, "lin.ckpt")
    torch.save({"state_dict": {"lin.weight": w.half(), "lin.bias": b}}, path)
    lin = Linear(8, 16)
    update_state(lin, load_weights(path)["state_dict"], "lin")
    assert torch.equal(lin.weight.cpu(), w.half().float()) and torch.equal(lin.bias.cpu(), b)


def test_load_weights_rejects_non_zip(tmp_path):
    from tinyfusers_b200.storage.unpicker import load_weights
    p = os.path.join(tmp_path, "x.bin")
    open(p, "wb").write(b"not a checkpoint")
    with pytest.raises(NameError):
        load_weights(p)


def _toy_vocab(tmp_path):
    from tinyfusers_b200.tokenizer.clip import bytes_to_unicode
    # merges file: header line, then one rule per line (the format of bpe_simple_vocab_16e6.txt.gz)
    rules = ["c a", "ca t</w>", "a t</w>", "h o", "ho r", "hor s", "hors e</w>", "b a", "ba g", "bag e", "bage l</w>"]
    path = os.path.join(tmp_path, "toy_bpe.txt.gz")
    with gzip.open(path, "wb") as fh:
        fh.write(("#version: toy\n" + "\n".join(rules) + "\n").encode())
    return path, rules, bytes_to_unicode()


def test_tokenizer_framing_and_merges(tmp_path):
    from tinyfusers_b200.tokenizer.clip import ClipTokenizer
    path, rules, b2u = _toy_vocab(tmp_path)
    tok = ClipTokenizer(path)
    sym = list(b2u.values())
    base = {s: i for i, s in enumerate(sym)}
    eow = {s + "</w>": 256 + i for i, s in enumerate(sym)}
    merged = {"".join(r.split()): 512 + i for i, r in enumerate(rules)}
    ids = tok.encode("  A  horse   cat bagel!  ")
    assert len(ids) == 77 and ids[0] == 49406 and ids[-1] == 49407
    body = ids[1:ids.index(49407)]
    # "a" -> a</w>; "horse" -> horse</w>; "cat" -> "ca"+"t</w>" merged to cat</w> (rule order); "bagel!" -> bage + l + !</w>
    assert body == [eow["a</w>"], merged["horse</w>"], merged["cat</w>"], merged["bage"], base["l"], eow["!</w>"]]
    assert tok.encode("") == [49406] + [49407] * 76
    long = tok.encode("cat " * 200)
    assert len(long) == 77 and long[0] == 49406 and long[-1] == 49407 and long[1:76] == [merged["cat</w>"]] * 75


def test_tokenizer_without_vocab_fails_loudly(monkeypatch):
    from tinyfusers_b200.tokenizer.clip import ClipTokenizer
    monkeypatch.delenv("TINYFUSERS_BPE_PATH", raising=False)
    with pytest.raises(RuntimeError, match="merges file"):
        ClipTokenizer()


def test_tokenizer_matches_reference_golden_ids(tmp_path):
    """Ids produced by the reference's OWN tokenizer (tests/golden/tokenizer_golden.json, oracle/make_tokenizer_golden.py imports
    /root/reference/tinyfusers/tokenizer/clip.py unmodified) on a merges file stored in the golden: ours must give the same 77
    integers for every prompt - lower-casing, whitespace cleaning, contractions, non-ASCII bytes, repeated symbols, in-text
    special tokens, truncation at 75, framing and padding. Integer work: bit-exact."""
    import gzip
    import json
    from tinyfusers_b200.tokenizer.clip import ClipTokenizer
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tokenizer_golden.json")))
    path = os.path.join(tmp_path, "learned_bpe.txt.gz")
    with gzip.open(path, "wb") as fh:
        fh.write(gold["merges"].encode("utf-8"))
    tok = ClipTokenizer(path)
    assert len(gold["cases"]) >= 10
    for case in gold["cases"]:
        assert "ids" in case, case
        assert tok.encode(case["prompt"]) == case["ids"], case["prompt"]
        assert len(case["ids"]) == 77 and case["ids"][0] == 49406 and case["ids"][-1] == 49407
