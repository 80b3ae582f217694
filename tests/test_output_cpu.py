"""Output side against a fixture produced by the reference's own tokenizer.decode /
predict_output_to_tsv_row / tsv_writer (oracle/make_output_golden.py)."""
import base64
import json
import os

import torch

from vitcap_b200 import output

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "out_rows.json")))


def _detok():
    return output.WordPieceDetokenizer(G["vocab"])


def test_rows_match_reference_rows():
    ids, lp = torch.tensor(G["ids"]), torch.tensor(G["logprobs"], dtype=torch.float32)
    rows = list(output.predict_output_to_tsv_rows(G["keys"], ids, lp, _detok()))
    assert [list(r) for r in rows] == G["rows"]


def test_tsv_files_byte_identical(tmp_path):
    ids, lp = torch.tensor(G["ids"]), torch.tensor(G["logprobs"], dtype=torch.float32)
    f = str(tmp_path / "sub" / "pred.tsv")
    output.tsv_writer(output.predict_output_to_tsv_rows(G["keys"], ids, lp, _detok()), f)
    for ext, b64 in G["files"].items():
        assert open(f[:-4] + ext, "rb").read() == base64.b64decode(b64), ext
    assert not os.path.exists(f + ".tmp")


def test_write_predictions_drops_duplicate_keys(tmp_path):
    ids, lp = torch.tensor(G["ids"]), torch.tensor(G["logprobs"], dtype=torch.float32)
    f = str(tmp_path / "pred.tsv")
    n = output.write_predictions(f, G["keys"], ids, lp, _detok())
    assert n == len(set(G["keys"])) == 5
    lines = open(f, "rb").read().decode().splitlines()
    assert [l.split("\t")[0] for l in lines] == ["img_000", "img_001", "img_002", "img_003", "img_005"]
    offs = [int(x) for x in open(f[:-4] + ".lineidx").read().split()]
    data = open(f, "rb").read()
    assert all(data[o:].startswith(l.encode()) for o, l in zip(offs, lines))
    b8 = open(f[:-4] + ".lineidx.8b", "rb").read()
    assert [int.from_bytes(b8[i:i + 8], "little") for i in range(0, len(b8), 8)] == offs


def test_detokenizer_from_token_list_and_empty_caption():
    d = output.WordPieceDetokenizer(["[PAD]", "a", "##b", "c", "[SEP]", "[CLS]"])
    assert d.decode([5, 1, 2, 3, 4, 0, 0]) == "ab c"
    assert d.decode([5, 4, 0]) == ""
    assert d.decode([5, 1, 99, 4]) == "a [UNK]"
