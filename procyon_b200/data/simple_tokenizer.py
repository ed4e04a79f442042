"""A dependency-free stand-in for the HF tokenizer surface UnifiedProCyon touches.

The reference builds its tokenizer from the Llama-3 files under $LLAMA3_PATH (model_unified.py:1088-1133), which do
not exist offline. Tests and the synthetic benchmark use this class instead; production passes a real
`transformers` tokenizer to `UnifiedProCyon(..., tokenizer=...)`. Only the members the model calls are provided:
`__call__`, `encode`, `add_tokens`, `batch_decode`, `__len__`, `pad/sep/eos/bos` tokens + ids, `padding_side`.
Words map to ids by a stable hash into [n_reserved, base_vocab); added special tokens get ids >= base_vocab, in
order of addition, exactly like `add_tokens` on a HF tokenizer.
"""
from __future__ import annotations

import re
import zlib
from typing import Dict, List, Union


class SimpleTokenizer:
    def __init__(self, base_vocab: int = 128256, bos_token: str = "<|begin_of_text|>", eos_token: str = "<|end_of_text|>"):
        self.base_vocab = base_vocab
        self.added: Dict[str, int] = {}
        self.id_to_added: Dict[int, str] = {}
        self.words: Dict[int, str] = {}
        self.bos_token, self.bos_token_id = bos_token, 0
        self.eos_token, self.eos_token_id = eos_token, 1
        self.pad_token = self.pad_token_id = None
        self.sep_token = self.sep_token_id = None
        self.padding_side = "right"
        self._n_reserved = 2

    def __len__(self):
        return self.base_vocab + len(self.added)

    def add_tokens(self, tok: Union[str, List[str]]):
        toks = [tok] if isinstance(tok, str) else tok
        for t in toks:
            if t not in self.added:
                idx = self.base_vocab + len(self.added)
                self.added[t] = idx
                self.id_to_added[idx] = t
        return len(toks)

    def _split(self, text: str) -> List[str]:
        if not self.added:
            return re.findall(r"\S+", text)
        pat = "(" + "|".join(re.escape(t) for t in sorted(self.added, key=len, reverse=True)) + ")"
        out = []
        for piece in re.split(pat, text):
            if piece in self.added:
                out.append(piece)
            else:
                out.extend(re.findall(r"\w+|[^\w\s]", piece))
        return out

    def _word_id(self, w: str) -> int:
        i = self._n_reserved + zlib.crc32(w.encode("utf-8")) % (self.base_vocab - self._n_reserved)
        self.words.setdefault(i, w)
        return i

    def encode(self, text: str, add_special_tokens: bool = True) -> List[int]:
        ids = [self.added[w] if w in self.added else self._word_id(w) for w in self._split(text)]
        return ([self.bos_token_id] + ids) if add_special_tokens else ids

    def __call__(self, text, padding=False, truncation=False, add_special_tokens=True, max_length=None, **kw):
        single = isinstance(text, str)
        texts = [text] if single else list(text)
        ids = [self.encode(t, add_special_tokens) for t in texts]
        if truncation and max_length is not None:
            ids = [x[:max_length] for x in ids]

        class _Enc(dict):
            __getattr__ = dict.__getitem__

        return _Enc(input_ids=ids[0] if single else ids)

    def decode(self, ids) -> str:
        out = []
        for i in [int(x) for x in ids]:
            if i == self.bos_token_id:
                out.append(self.bos_token)
            elif i == self.eos_token_id:
                out.append(self.eos_token)
            elif i in self.id_to_added:
                out.append(self.id_to_added[i])
            else:
                out.append(self.words.get(i, f"<{i}>"))
        return " ".join(out)

    def batch_decode(self, batch, **kw) -> List[str]:
        return [self.decode(row.tolist() if hasattr(row, "tolist") else row) for row in batch]
