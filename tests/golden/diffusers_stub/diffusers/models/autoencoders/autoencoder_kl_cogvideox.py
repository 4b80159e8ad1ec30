class AutoencoderKLCogVideoX:
    """Name only (isinstance check at /root/reference/src/dwm/pipelines/ctsd.py:963-964)."""
