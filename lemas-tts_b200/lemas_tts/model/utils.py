"""Host helpers of the sampler (reference lemas_tts/model/utils.py): seeding, masks, tokenizer.

Only what the inference path touches is provided; the pinyin conversion of the reference
imports a module that does not exist in its own tree and is dead on the `TTS.infer` path.
"""
from __future__ import annotations

import os
import random

import torch
from torch.nn.utils.rnn import pad_sequence


def seed_everything(seed=0):
    """model/utils.py:18-25."""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def exists(v):
    return v is not None


def default(v, d):
    return v if exists(v) else d


def lens_to_mask(t: torch.Tensor, length: int | None = None) -> torch.Tensor:
    """model/utils.py:42-47: [b] lengths -> bool [b, n]."""
    if length is None:
        length = int(t.amax())
    return torch.arange(length, device=t.device)[None, :] < t[:, None]


def list_str_to_tensor(text: list[str], padding_value=-1) -> torch.Tensor:
    """model/utils.py:80-83 (UTF-8 byte tokenizer)."""
    return pad_sequence([torch.tensor([*bytes(t, "UTF-8")]) for t in text], padding_value=padding_value,
                        batch_first=True)


def list_str_to_idx(text, vocab_char_map: dict[str, int], padding_value=-1) -> torch.Tensor:
    """model/utils.py:86-94: tokens -> ids, unknown -> 0, rows padded with -1."""
    rows = [torch.tensor([vocab_char_map.get(c, 0) for c in t], dtype=torch.long) for t in text]
    return pad_sequence(rows, padding_value=padding_value, batch_first=True)


def get_tokenizer(dataset_name, tokenizer: str = "pinyin"):
    """model/utils.py:98-128.  "custom": dataset_name is the path of a vocab.txt (one token per line)."""
    if tokenizer == "byte":
        return None, 256
    if tokenizer in ("pinyin", "char"):
        from importlib.resources import files

        path = os.path.join(files("lemas_tts").joinpath("../../data"), f"{dataset_name}_{tokenizer}/vocab.txt")
    elif tokenizer == "custom":
        path = dataset_name
    else:
        raise ValueError(f"unknown tokenizer {tokenizer!r}")
    vocab_char_map = {}
    with open(path, "r", encoding="utf-8") as f:
        for i, char in enumerate(f):
            vocab_char_map[char[:-1]] = i
    if tokenizer in ("pinyin", "char"):
        assert vocab_char_map[" "] == 0, "make sure space is of idx 0 in vocab.txt, cuz 0 is used for unknown char"
    return vocab_char_map, len(vocab_char_map)


def convert_char_to_pinyin(text_list, polyphone=True):
    """model/utils.py:132-176 needs jieba/pypinyin and a module the reference itself does not ship
    (lemas_tts.infer.cn_tn); the phone frontend never reaches it (ref_text is a list there)."""
    raise ImportError("convert_char_to_pinyin: the raw-string path needs jieba/pypinyin, which this build "
                      "does not bundle; pass phone lists (TTS.infer does)")
