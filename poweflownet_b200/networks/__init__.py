from .MPN import EdgeAggregation, MaskEmbdMultiMPN, TAGConv  # noqa: F401
