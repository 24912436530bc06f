"""Utilities (reference ``stream_sim/utils.py``)."""
import yaml


def parse_config(config):
    """Parse a YAML file path, or a YAML string, into a dict (reference utils.py:7-23)."""
    try:
        with open(config, "r") as handle:
            return yaml.safe_load(handle)
    except (OSError, FileNotFoundError):
        return yaml.safe_load(config)
