"""Import shim (test infrastructure only): `timm` is absent from this image and the
reference imports two names from it.  Neither does inference arithmetic."""
