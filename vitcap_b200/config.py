"""Static description of the ViTCAP captioning path.

Mirrors what the reference spreads over ``BertConfig`` (yaml/VILT-L12-H784-uncased_*/config.json),
the pipeline's ``get_fusion_config`` (tagger_caption_uni_pipeline_expanding_bertemb.py:520-564)
and the timm registry entries ``vit_base_patch16_{224,384}`` / ``vit_base_patch32_384``
(vision_transformer.py:1195-1278).
"""
from dataclasses import dataclass, field, asdict


@dataclass(frozen=True)
class VitCapConfig:
    img_size: int = 384
    patch: int = 16
    hidden: int = 768
    heads: int = 12
    inter: int = 3072
    vocab: int = 30522
    enc_blocks: int = 12          # model.bert.encoder.blocks
    split_blocks: int = 4         # model.bert.encoder.tag_blocks (copy of the last 4)
    dec_layers: int = 4           # modeling_bert.py:1342-1346 (decoder_layer or 4)
    max_pos: int = 512
    type_vocab: int = 2
    topk: int = 50                # concept tags kept (yaml: topk)
    tag_thresh: float = 0.2       # modeling_bert.py:1432
    vit_ln_eps: float = 1e-6      # vision_transformer.py:352
    bert_ln_eps: float = 1e-12    # config.json layer_norm_eps
    max_seq_a: int = 20           # caption slots (max_seq_a_length)
    max_seq: int = 70             # caption + od/tag slots (max_seq_length)
    sep_id: int = 102             # [SEP] of bert-base-uncased: forced into the last label slot (modeling_bert.py:1447)
    # label-slot embedding recipe (modeling_bert.py:1454 / 1480). 'cls' is what the shipped yaml sets and what is implemented;
    # the pipeline's fallback 'bert' (file line 554: encode_tag_to_embedding(cls_emb=None) / extra_embeddings) only matters when
    # a label region is VISIBLE, and is refused there (model.py _label_counts)
    tagemb: str = "cls"

    @property
    def head_dim(self):
        return self.hidden // self.heads

    @property
    def grid(self):
        return self.img_size // self.patch

    @property
    def n_patches(self):
        return self.grid * self.grid

    @property
    def n_tokens(self):          # N: cls + patches
        return self.n_patches + 1

    @property
    def n_ctx(self):             # C: tag-CLS + image tokens (modeling_bert.py:1493)
        return self.n_tokens + 1

    @property
    def patch_dim(self):
        return 3 * self.patch * self.patch

    def to_dict(self):
        return asdict(self)


VARIANTS = {
    "16_384": VitCapConfig(img_size=384, patch=16),
    "16_224": VitCapConfig(img_size=224, patch=16),
    "32_384": VitCapConfig(img_size=384, patch=32),
}


def variant(name, **overrides):
    base = VARIANTS[name]
    d = base.to_dict()
    d.update(overrides)
    return VitCapConfig(**d)


def tiny(**overrides):
    """A structurally identical but small model for fast kernel/host tests
    (hidden and head_dim keep the production values the kernels are tuned for)."""
    d = dict(img_size=64, patch=16, enc_blocks=3, split_blocks=1, dec_layers=2, vocab=3000, inter=1536)
    d.update(overrides)
    return VitCapConfig(**d)
