"""Text front end: hypothesis string -> wordpiece ids.

The reference calls ``BertTokenizer.from_pretrained("bert-base-chinese").tokenize``
(MLM_PLL/preprocess.py:10,34).  That vocabulary is fetched from the HF hub at run time
and is not available offline, so two modes exist:

* ``BertCharTokenizer(vocab_file)``: BERT's BasicTokenizer + WordPiece restated for the
  case that matters here — every CJK character is its own token; runs of other
  characters are lower-cased, accent-stripped, split on punctuation and greedily
  longest-match wordpieced; unknown -> [UNK].  (transformers
  models/bert/tokenization_bert.py; checked against it in tests/test_host.py.)
* ``SyntheticCharTokenizer()``: deterministic char -> id map used with random-init
  weights (synth.synthetic_token_id).
"""
from __future__ import annotations

import unicodedata
from typing import Dict, List

from .synth import UNK_ID, synthetic_token_id


def _is_cjk(cp: int) -> bool:
    return (0x4E00 <= cp <= 0x9FFF or 0x3400 <= cp <= 0x4DBF or 0x20000 <= cp <= 0x2A6DF or
            0x2A700 <= cp <= 0x2B73F or 0x2B740 <= cp <= 0x2B81F or 0x2B820 <= cp <= 0x2CEAF or
            0xF900 <= cp <= 0xFAFF or 0x2F800 <= cp <= 0x2FA1F)


def _is_punct(ch: str) -> bool:
    cp = ord(ch)
    if 33 <= cp <= 47 or 58 <= cp <= 64 or 91 <= cp <= 96 or 123 <= cp <= 126:
        return True
    return unicodedata.category(ch).startswith("P")


def _is_control(ch: str) -> bool:
    if ch in ("\t", "\n", "\r"):
        return False
    return unicodedata.category(ch).startswith("C")


def _is_whitespace(ch: str) -> bool:
    return ch in (" ", "\t", "\n", "\r") or unicodedata.category(ch) == "Zs"


class SyntheticCharTokenizer:
    def tokenize(self, text: str) -> List[str]:
        return list(text)

    def convert_tokens_to_ids(self, tokens: List[str]) -> List[int]:
        return [synthetic_token_id(t) for t in tokens]

    def encode(self, text: str) -> List[int]:
        return self.convert_tokens_to_ids(self.tokenize(text))


class BertCharTokenizer:
    def __init__(self, vocab_file: str, do_lower_case: bool = True, unk_token: str = "[UNK]",
                 max_input_chars_per_word: int = 100):
        self.vocab: Dict[str, int] = {}
        with open(vocab_file, "r", encoding="utf-8") as f:
            for i, line in enumerate(f):
                self.vocab[line.rstrip("\n")] = i
        self.do_lower_case = do_lower_case
        self.unk_token = unk_token
        self.max_chars = max_input_chars_per_word

    # BasicTokenizer.tokenize
    def _basic(self, text: str) -> List[str]:
        out = []
        for ch in text:
            cp = ord(ch)
            if cp == 0 or cp == 0xFFFD or _is_control(ch):
                continue
            out.append(" " if _is_whitespace(ch) else ch)
        text = unicodedata.normalize("NFC", "".join(out))
        spaced = []
        for ch in text:
            spaced.append(f" {ch} " if _is_cjk(ord(ch)) else ch)
        words = []
        for tok in "".join(spaced).strip().split():
            if self.do_lower_case:
                tok = tok.lower()
                tok = "".join(c for c in unicodedata.normalize("NFD", tok) if unicodedata.category(c) != "Mn")
            cur = []
            for ch in tok:
                if _is_punct(ch):
                    if cur:
                        words.append("".join(cur))
                        cur = []
                    words.append(ch)
                else:
                    cur.append(ch)
            if cur:
                words.append("".join(cur))
        return " ".join(words).split()

    # WordpieceTokenizer.tokenize
    def _wordpiece(self, word: str) -> List[str]:
        if len(word) > self.max_chars:
            return [self.unk_token]
        pieces, start = [], 0
        while start < len(word):
            end, cur = len(word), None
            while start < end:
                sub = word[start:end]
                if start > 0:
                    sub = "##" + sub
                if sub in self.vocab:
                    cur = sub
                    break
                end -= 1
            if cur is None:
                return [self.unk_token]
            pieces.append(cur)
            start = end
        return pieces

    def tokenize(self, text: str) -> List[str]:
        return [p for w in self._basic(text) for p in self._wordpiece(w)]

    def convert_tokens_to_ids(self, tokens: List[str]) -> List[int]:
        unk = self.vocab.get(self.unk_token, UNK_ID)
        return [self.vocab.get(t, unk) for t in tokens]

    def encode(self, text: str) -> List[int]:
        return self.convert_tokens_to_ids(self.tokenize(text))
