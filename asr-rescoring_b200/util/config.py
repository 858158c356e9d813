"""Nested-dict -> nested SimpleNamespace, as util/config.py:3-15 of the reference."""
from types import SimpleNamespace


def parse_config(config: dict) -> SimpleNamespace:
    ns = SimpleNamespace()
    for key, value in config.items():
        setattr(ns, key, parse_config(value) if isinstance(value, dict) else value)
    return ns
