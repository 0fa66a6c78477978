"""Tokenizers of the Flux text path (reference: flux/tokenizers.py).  Integer path -> bit-exact.

Host-only Python with the reference's behaviour: the CLIP tokenizer lower-cases, collapses
whitespace, splits with the CLIP regex, applies greedy lowest-rank BPE merges, wraps in BOS/EOS and
truncates to 77 keeping EOS last; a single prompt is NOT padded (flux/tokenizers.py:110-119).  The
T5 tokenizer is SentencePiece + EOS, padded with id 0 to 256/512 and never truncated
(flux/tokenizers.py:160-185).  Both return int32 torch tensors [B, n].
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import regex
import torch

_CLIP_PATTERN = r"""<\|startoftext\|>|<\|endoftext\|>|'s|'t|'re|'ve|'m|'ll|'d|[\p{L}]+|[\p{N}]|[^\s\p{L}\p{N}]+"""


class CLIPTokenizer:
    bos = "<|startoftext|>"
    eos = "<|endoftext|>"

    def __init__(self, bpe_ranks: Dict[Tuple[str, str], int], vocab: Dict[str, int], max_length: int = 77):
        self.max_length = max_length
        self.bpe_ranks = bpe_ranks
        self.vocab = vocab
        self.pat = regex.compile(_CLIP_PATTERN, regex.IGNORECASE)
        self._cache = {self.bos: self.bos, self.eos: self.eos}

    @property
    def bos_token(self) -> int:
        return self.vocab[self.bos]

    @property
    def eos_token(self) -> int:
        return self.vocab[self.eos]

    def bpe(self, word: str):
        """Greedy merge of the best-ranked adjacent pair until none is in the merge table."""
        hit = self._cache.get(word)
        if hit is not None:
            return hit
        parts: List[str] = list(word[:-1]) + [word[-1] + "</w>"]
        inf = float("inf")
        while len(parts) > 1:
            pairs = set(zip(parts, parts[1:]))
            best = min(pairs, key=lambda pr: self.bpe_ranks.get(pr, inf))
            if best not in self.bpe_ranks:
                break
            merged, i = [], 0
            while i < len(parts):
                if i + 1 < len(parts) and (parts[i], parts[i + 1]) == best:
                    merged.append(parts[i] + parts[i + 1])
                    i += 2
                else:
                    merged.append(parts[i])
                    i += 1
            parts = merged
        if len(word) > 1 or True:
            self._cache[word] = parts
        return parts

    def tokenize(self, text, prepend_bos: bool = True, append_eos: bool = True):
        if isinstance(text, list):
            return [self.tokenize(t, prepend_bos, append_eos) for t in text]
        clean = regex.sub(r"\s+", " ", text.lower())
        ids = [self.vocab[piece] for w in regex.findall(self.pat, clean) for piece in self.bpe(w)]
        if prepend_bos:
            ids = [self.bos_token] + ids
        if append_eos:
            ids.append(self.eos_token)
        if len(ids) > self.max_length:
            ids = ids[: self.max_length]
            if append_eos:
                ids[-1] = self.eos_token
        return ids

    def encode(self, text) -> torch.Tensor:
        if not isinstance(text, list):
            return self.encode([text])
        rows = self.tokenize(text)
        n = max(len(r) for r in rows)
        return torch.tensor([r + [self.eos_token] * (n - len(r)) for r in rows], dtype=torch.int32)


class T5Tokenizer:
    def __init__(self, model_file, max_length: int = 512):
        from sentencepiece import SentencePieceProcessor
        self._tokenizer = SentencePieceProcessor(model_file)
        self.max_length = max_length

    def _piece(self, idx):
        try:
            return self._tokenizer.id_to_piece(idx)
        except IndexError:
            return None

    @property
    def pad_token(self) -> int:
        return self._tokenizer.pad_id()

    @property
    def bos_token(self) -> int:
        return self._tokenizer.bos_id()

    @property
    def eos_token(self) -> int:
        return self._tokenizer.eos_id()

    pad = property(lambda self: self._piece(self.pad_token))
    bos = property(lambda self: self._piece(self.bos_token))
    eos = property(lambda self: self._piece(self.eos_token))

    def tokenize(self, text, prepend_bos: bool = True, append_eos: bool = True, pad: bool = True):
        if isinstance(text, list):
            return [self.tokenize(t, prepend_bos, append_eos, pad) for t in text]
        ids = list(self._tokenizer.encode(text))
        if prepend_bos and self.bos_token >= 0:
            ids = [self.bos_token] + ids
        if append_eos and self.eos_token >= 0:
            ids.append(self.eos_token)
        if pad and len(ids) < self.max_length and self.pad_token >= 0:
            ids += [self.pad_token] * (self.max_length - len(ids))
        return ids

    def encode(self, text, pad: bool = True) -> torch.Tensor:
        if not isinstance(text, list):
            return self.encode([text], pad=pad)
        fill = self.pad_token if self.pad_token >= 0 else 0
        rows = self.tokenize(text, pad=pad)
        n = max(len(r) for r in rows)
        return torch.tensor([r + [fill] * (n - len(r)) for r in rows], dtype=torch.int32)


class SyntheticTokenizer:
    """Stand-in used when no tokenizer files exist offline: maps a prompt string to deterministic
    pseudo-token ids shaped like the real tokenizer's output (seeded by a CRC of the text)."""

    def __init__(self, kind: str, max_length: int, vocab: int):
        self.kind, self.max_length, self.vocab = kind, max_length, vocab

    def encode(self, text, pad: bool = True) -> torch.Tensor:
        import zlib
        from .synthetic import synthetic_prompt_tokens
        if isinstance(text, list):
            return torch.cat([self.encode(t, pad) for t in text], 0)
        seed = zlib.crc32(text.encode()) & 0x7FFFFFFF
        n_tok = max(1, min(len(text.split()) * 2, 60, self.max_length - 2))
        if self.kind == "t5":
            return synthetic_prompt_tokens(self.max_length, 77, seed, n_tok, t5_vocab=self.vocab, pad=pad)[0]
        return synthetic_prompt_tokens(self.max_length, 77, seed, n_tok, clip_vocab=self.vocab)[1]
