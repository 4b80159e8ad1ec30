from oracle.d31 import AdapterBlock, AdapterResnetBlock  # noqa: F401
