from . import adapter, autoencoders, attention, attention_processor, embeddings, resnet, transformers, unets  # noqa: F401
