from . import adapter, attention, attention_processor, embeddings, resnet, transformers, unets  # noqa: F401
