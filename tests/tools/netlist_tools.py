"""Fixture tooling: the blueprint loader lives in the product (iyokan_b200/blueprint.py); here it is pointed at the
REFERENCE's pre-synthesised mux-ram netlists so that derived fixtures have the reference's exact gate counts
(runs where /root/reference exists; see tests/golden/make_netlists.py)."""
from iyokan_b200.blueprint import *  # noqa: F401,F403
from iyokan_b200 import blueprint as _bp


def read_blueprint(toml_path, reference_src="/root/reference/src"):
    return _bp.read_blueprint(toml_path, mux_ram_json_dir=reference_src)
