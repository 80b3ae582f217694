"""Output side of the caption path: token ids -> caption strings -> prediction TSV.

Restates, for the rows the hot path produces,
  * ``CaptionUniPipeline.predict_output_to_tsv_row`` (tagger_caption_uni_pipeline_expanding_bertemb.py:620-630):
    ``key \\t json([{'caption': str, 'conf': exp(logprob)}, ...])`` per image,
  * ``BertTokenizer.decode(ids, skip_special_tokens=True)`` (tokenization_utils.py:430-470, 506-510;
    tokenization_bert.py:184-191): drop special ids, join with ' ', merge ' ##' word pieces, clean up punctuation spacing,
  * ``tsv_writer`` (tsv_io.py:959-997): the TSV plus its ``.lineidx`` (decimal byte offsets) and ``.lineidx.8b``
    (little-endian uint64 offsets) companions, written to ``.tmp`` names and renamed.
With the packed all-gather of ``vitcap_b200.parallel`` rank 0 holds every image's result in dataset order, so the
reference's per-rank files + concat + reorder + de-duplicate step (uni_pipeline.py:813-831) reduces to ``write_predictions``.
"""
import json
import os

import torch

SPECIAL_TOKENS = ("[UNK]", "[SEP]", "[PAD]", "[CLS]", "[MASK]")


class WordPieceDetokenizer:
    """ids -> text with the reference tokenizer's ``decode`` semantics. ``vocab`` is a vocab.txt path (one token per line,
    line number = id), a list of tokens, or a {id: token} dict (ids missing from a dict decode to [UNK])."""

    def __init__(self, vocab, unk_token="[UNK]"):
        if isinstance(vocab, str):
            with open(vocab, "r", encoding="utf-8") as f:
                toks = [line.rstrip("\n") for line in f]
            self.ids_to_tokens = dict(enumerate(toks))
        elif isinstance(vocab, dict):
            self.ids_to_tokens = {int(k): v for k, v in vocab.items()}
        else:
            self.ids_to_tokens = dict(enumerate(vocab))
        self.unk_token = unk_token
        tok_to_id = {}
        for i, t in self.ids_to_tokens.items():
            tok_to_id.setdefault(t, i)
        # all_special_ids (tokenization_utils.py:497-503); BERT ids: [PAD] 0, [UNK] 100, [CLS] 101, [SEP] 102, [MASK] 103
        self.special_ids = {tok_to_id[t] for t in SPECIAL_TOKENS if t in tok_to_id}

    def decode(self, token_ids, skip_special_tokens=True, clean_up_tokenization_spaces=True):
        toks = []
        for i in token_ids:
            i = int(i)
            if skip_special_tokens and i in self.special_ids:
                continue
            toks.append(self.ids_to_tokens.get(i, self.unk_token))
        text = " ".join(toks).replace(" ##", "").strip()
        if clean_up_tokenization_spaces:
            text = clean_up_tokenization(text)
        return text


def clean_up_tokenization(s):
    """tokenization_utils.py:506-510."""
    return (s.replace(" .", ".").replace(" ?", "?").replace(" !", "!").replace(" ,", ",").replace(" ' ", "'")
            .replace(" n't", "n't").replace(" 'm", "'m").replace(" do not", " don't").replace(" 's", "'s")
            .replace(" 've", "'ve").replace(" 're", "'re"))


def predict_output_to_tsv_rows(keys, ids, logprobs, detok):
    """Yields ``(key, json)`` exactly like the reference row generator. ids (B, keep, L) int, logprobs (B, keep) fp32."""
    confs = torch.exp(logprobs)                           # on whatever device the caller holds them, as the reference does
    ids_l = ids.tolist()
    confs_l = confs.tolist()                              # == [c.item() for c in row]: fp32 values widened to Python floats
    for key, caps, cf in zip(keys, ids_l, confs_l):
        res = [{"caption": detok.decode(cap, skip_special_tokens=True), "conf": c} for cap, c in zip(caps, cf)]
        yield key, json.dumps(res)


def tsv_writer(values, tsv_file_name, sep="\t"):
    """Row iterator -> ``name.tsv`` + ``name.lineidx`` + ``name.lineidx.8b`` (tsv_io.py:959-997, Python-3 branch)."""
    d = os.path.dirname(tsv_file_name)
    if d:
        os.makedirs(d, exist_ok=True)
    lineidx = os.path.splitext(tsv_file_name)[0] + ".lineidx"
    idx8b = lineidx + ".8b"
    sepb = sep.encode()
    off = 0
    with open(tsv_file_name + ".tmp", "wb") as fp, open(lineidx + ".tmp", "w") as fpidx, open(idx8b + ".tmp", "wb") as fp8:
        for value in values:
            assert value is not None
            v = sepb.join(x if type(x) == bytes else str(x).encode() for x in value) + b"\n"
            fp.write(v)
            fpidx.write(str(off) + "\n")
            fp8.write(off.to_bytes(8, "little"))
            off += len(v)
    os.rename(tsv_file_name + ".tmp", tsv_file_name)
    os.rename(lineidx + ".tmp", lineidx)
    os.rename(idx8b + ".tmp", idx8b)


def write_predictions(predict_file, keys, ids, logprobs, detok):
    """Writes the merged prediction file from gathered results in dataset order; rows whose key was already written (the
    wrap-around padding of the contiguous-chunk sampler) are dropped, as uni_pipeline.py:822-828 does by key."""
    seen = set()

    def rows():
        for key, js in predict_output_to_tsv_rows(keys, ids, logprobs, detok):
            if key in seen:
                continue
            seen.add(key)
            yield key, js

    tsv_writer(rows(), predict_file)
    return len(seen)
