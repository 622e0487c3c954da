"""Shim: the reference imports pyplot only for debug plots."""
