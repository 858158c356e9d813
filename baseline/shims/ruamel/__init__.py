"""Stand-in for the `ruamel` namespace package (not installed in this image)."""
