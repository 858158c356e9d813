"""`--config file.yaml` -> SimpleNamespace; same surface as util/arg_parser.py:6-22 of the
reference (which uses ruamel.yaml's `yaml.load(..., Loader=yaml.Loader)`; PyYAML's call is
API-compatible and is what this image has).  A missing key is an AttributeError at first
use, exactly like the reference — there are no defaults and no validation."""
import argparse

try:  # the reference imports ruamel.yaml; fall back to PyYAML (same call signature)
    import ruamel.yaml as yaml  # type: ignore
    _ = yaml.Loader
except Exception:  # pragma: no cover - depends on the environment
    import yaml

from .config import parse_config


class ArgParser():
    def __init__(self) -> None:
        self.parser = argparse.ArgumentParser()
        self.parser.add_argument("--config", type=str, required=True, help="yaml configuration file path")

    def parse(self, argv=None):
        args = self.parser.parse_args(argv)
        with open(args.config, "r") as f:
            config = yaml.load(f, Loader=yaml.Loader)
        return parse_config(config)
