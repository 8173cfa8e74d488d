"""CLIP byte-level BPE tokenizer with the reference's interface (reference: tinyfusers/tokenizer/clip.py:10-75):
`ClipTokenizer(bpe_path).encode(text)` -> 77 ids, BOS 49406, padded with EOS 49407.

Pure host-side glue. The reference downloads `bpe_simple_vocab_16e6.txt.gz` at import
time (clip.py:7-8); there is no network here, so the merges file must be given (argument, or the environment
variable TINYFUSERS_BPE_PATH) and a clear error is raised without it. The vocabulary layout is the one the file
format defines: 256 byte symbols, the same 256 with the end-of-word marker, one entry per merge rule, then the two
special tokens."""
import gzip
import os
import re

BOS, EOS, CONTEXT = 49406, 49407, 77
_N_MERGES = 49152 - 256 - 2
_EOW = "</w>"


def default_bpe():
    path = os.environ.get("TINYFUSERS_BPE_PATH")
    if not path or not os.path.exists(path):
        raise RuntimeError("ClipTokenizer needs the CLIP merges file bpe_simple_vocab_16e6.txt.gz: pass bpe_path or set "
                           "TINYFUSERS_BPE_PATH (the reference downloads it; this environment has no network)")
    return path


def bytes_to_unicode():
    """byte value -> printable unicode character: printable latin-1 bytes map to themselves, the 68 others to
    code points from 256 upwards, in byte order (the byte-level BPE convention the vocabulary file was built with)."""
    keep = set(range(ord("!"), ord("~") + 1)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    ordered = sorted(keep, key=lambda b: (0 if b <= ord("~") else 1 if b <= 0xAC else 2, b))
    table = {b: chr(b) for b in ordered}
    extra = 0
    for b in range(256):
        if b not in keep:
            table[b] = chr(256 + extra)
            extra += 1
    return table


def whitespace_clean(text):
    return re.sub(r"\s+", " ", text).strip()


def get_pairs(word):
    """Set of adjacent symbol pairs of a word given as a tuple of symbols."""
    return set(zip(word[:-1], word[1:]))


class ClipTokenizer:
    def __init__(self, bpe_path: str = None):
        bpe_path = default_bpe() if bpe_path is None else bpe_path
        opener = gzip.open if str(bpe_path).endswith(".gz") else open
        with opener(bpe_path, "rb") as fh:
            lines = fh.read().decode("utf-8").split("\n")
        # line 0 is the file's version header. Every line of the slice takes a vocabulary slot, as in the reference
        # (clip.py:13-19) - including an empty last line when the file is shorter than the slice - so ids agree for any file
        rules = [tuple(l.split()) for l in lines[1:_N_MERGES + 1]]
        self.byte_encoder = bytes_to_unicode()
        symbols = list(self.byte_encoder.values())
        # the two special tokens close the vocabulary (reference clip.py:20): with the full 48 894-rule file they land on
        # 49406 / 49407, the ids `encode` frames every prompt with; with a shorter file an in-text "<|endoftext|>" gets the
        # position the file implies, exactly as in the reference
        vocab = symbols + [s + _EOW for s in symbols] + ["".join(r) for r in rules] + ["<|startoftext|>", "<|endoftext|>"]
        self.encoder = {tok: i for i, tok in enumerate(vocab)}
        self.bpe_ranks = {r: i for i, r in enumerate(rules) if len(r) == 2}
        self.cache = {"<|startoftext|>": "<|startoftext|>", "<|endoftext|>": "<|endoftext|>"}
        self.pat = re.compile(r"<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[^\s]+", re.IGNORECASE)

    def bpe(self, token):
        """Applies the merge rules, lowest rank first, to one pre-token; returns the symbols joined by spaces."""
        hit = self.cache.get(token)
        if hit is not None:
            return hit
        word = list(token[:-1]) + [token[-1] + _EOW]
        while len(word) > 1:
            best = min(((self.bpe_ranks[p], p) for p in zip(word[:-1], word[1:]) if p in self.bpe_ranks), default=None)
            if best is None:
                break
            a, b = best[1]
            merged, i = [], 0
            while i < len(word):                      # merge every non-overlapping occurrence, left to right
                if i + 1 < len(word) and word[i] == a and word[i + 1] == b:
                    merged.append(a + b)
                    i += 2
                else:
                    merged.append(word[i])
                    i += 1
            word = merged
        out = " ".join(word)
        self.cache[token] = out
        return out

    def encode(self, text):
        ids = []
        for tok in self.pat.findall(whitespace_clean(text.strip()).lower()):
            tok = "".join(self.byte_encoder[b] for b in tok.encode("utf-8"))
            ids.extend(self.encoder[s] for s in self.bpe(tok).split(" "))
        ids = ids[:CONTEXT - 2]                      # keep two slots for the start / end tokens
        return [BOS] + ids + [EOS] * (CONTEXT - len(ids) - 1)
