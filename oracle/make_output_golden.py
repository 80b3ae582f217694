"""TEST INFRASTRUCTURE ONLY. Generates tests/golden/out_rows.json by running the UNMODIFIED reference's output code
(BertTokenizer.decode, CaptionUniPipeline.predict_output_to_tsv_row's row format, tsv_writer) on synthetic token ids.

    python -m oracle.make_output_golden          # in the build container (needs /root/reference)

The fixture embeds only the vocabulary entries it uses (id -> token), the ids / log-probs / keys, the expected
(key, json) rows and the bytes of the three files tsv_writer produces.
"""
import base64
import json
import os
import tempfile

import torch

from oracle import ref_loader

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "out_rows.json")


def main():
    ref_loader.install_shims()
    from src.layers.bert import BertTokenizer
    from src.tools.tsv.tsv_io import tsv_writer
    vocab_dir = os.path.join(ref_loader.REF_ROOT, "yaml", "VILT-L12-H784-uncased_16_384")
    tok = BertTokenizer.from_pretrained(vocab_dir, do_lower_case=True)
    texts = ["a cat walking on a beach near a body of water", "two people don't know what's in the child's hand , do they ?",
             "a man riding a skateboard ! it 's n't easy .", "an unbelievably photogenic giraffe , i 've seen", ""]
    g = torch.Generator().manual_seed(0)
    B, keep, L = 6, 2, 20
    ids = torch.zeros(B, keep, L, dtype=torch.int64)
    for b in range(B):
        for k in range(keep):
            t = texts[(b + k) % len(texts)]
            w = tok.convert_tokens_to_ids(tok.tokenize(t))[: L - 2]
            row = [101] + w + [102]
            ids[b, k, :len(row)] = torch.tensor(row)
    ids[5, 1, 3] = 100                                     # an [UNK] inside
    ids[4, 0, 5] = 103                                     # a stray [MASK]
    ids[3, 1, 2] = 30521                                   # last vocabulary id
    logprobs = -torch.rand(B, keep, generator=g) * 3
    logprobs[2, 1] = -1e5                                  # an unfilled beam slot (modeling_utils.py:1075)
    keys = ["img_%03d" % i for i in range(B)]
    keys[4] = keys[1]                                      # duplicate key (sampler wrap-around)

    def rows():                                            # predict_output_to_tsv_row, pipeline file lines 620-630
        all_caps, all_confs = ids, torch.exp(logprobs)
        for img_key, caps, confs in zip(keys, all_caps, all_confs):
            res = []
            for cap, conf in zip(caps, confs):
                cap = tok.decode(cap.tolist(), skip_special_tokens=True)
                res.append({'caption': cap, 'conf': conf.item()})
            yield img_key, json.dumps(res)

    expected = list(rows())
    with tempfile.TemporaryDirectory() as d:
        f = os.path.join(d, "pred.tsv")
        tsv_writer(rows(), f)
        files = {ext: base64.b64encode(open(os.path.join(d, "pred" + ext), "rb").read()).decode()
                 for ext in (".tsv", ".lineidx", ".lineidx.8b")}
    used = sorted(set(ids.flatten().tolist()) | {0, 100, 101, 102, 103})
    vocab = {str(i): tok.ids_to_tokens[i] for i in used}
    json.dump({"ids": ids.tolist(), "logprobs": logprobs.tolist(), "keys": keys, "rows": expected, "files": files, "vocab": vocab},
              open(OUT, "w"), indent=0)
    print("wrote", OUT, len(expected), "rows")


if __name__ == "__main__":
    main()
