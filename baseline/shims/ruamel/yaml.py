"""Stand-in for `ruamel.yaml` as used by the reference's util/config.py (load + Loader): PyYAML."""
from yaml import *  # noqa: F401,F403
from yaml import Loader, load  # noqa: F401
