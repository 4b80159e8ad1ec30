"""Import target only (PyAV is not installed; the preview writer is never called)."""
