"""Import target only (UNet family)."""
