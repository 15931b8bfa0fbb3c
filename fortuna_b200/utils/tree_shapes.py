"""Shape-only parameter pytrees for benchmarks and tests (no model code: the backbone is outside the hot path)."""


def bert_base_shapes(n_labels=2):
    """Leaf shapes of a BERT-base classifier (config 5), ~110 M parameters, ~200 leaves."""
    H, I, V, P = 768, 3072, 30522, 512
    shapes = {
        "embeddings": {
            "word_embeddings": {"embedding": (V, H)},
            "position_embeddings": {"embedding": (P, H)},
            "token_type_embeddings": {"embedding": (2, H)},
            "LayerNorm": {"scale": (H,), "bias": (H,)},
        },
        "pooler": {"dense": {"kernel": (H, H), "bias": (H,)}},
        "classifier": {"kernel": (H, n_labels), "bias": (n_labels,)},
    }
    enc = {}
    for i in range(12):
        enc[str(i)] = {
            "attention": {
                "self": {n: {"kernel": (H, H), "bias": (H,)} for n in ("query", "key", "value")},
                "output": {"dense": {"kernel": (H, H), "bias": (H,)}, "LayerNorm": {"scale": (H,), "bias": (H,)}},
            },
            "intermediate": {"dense": {"kernel": (H, I), "bias": (I,)}},
            "output": {"dense": {"kernel": (I, H), "bias": (H,)}, "LayerNorm": {"scale": (H,), "bias": (H,)}},
        }
    shapes["encoder"] = {"layer": enc}
    return {"model": {"params": shapes}}
