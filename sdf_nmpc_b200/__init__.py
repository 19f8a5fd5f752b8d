"""sdf-nmpc hot path, B200-native.  Directory helpers mirror ``sdf_nmpc/__init__.py:6-20``: the reference looks its
YAML files and its TorchScript weights (``cfg.nn.sdf_weights``, ``cfg.nn.vae_weights``) up in the package's ``config`` and
``data`` directories; ``set_data_dir`` points the lookup elsewhere (the reference's ``.pt`` files are Git-LFS objects that
a deployment drops into a directory of its choice)."""
import os
from pathlib import Path

_HERE = Path(os.path.dirname(os.path.abspath(__file__)))
_data_dir = None


def default_config_dir() -> Path:
    return _HERE / 'config'


def default_data_dir() -> Path:
    if _data_dir is not None:
        return Path(_data_dir)
    return Path(os.environ.get('SDFNMPC_DATA_DIR', _HERE / 'data'))


def set_data_dir(path):
    global _data_dir
    _data_dir = None if path is None else str(path)
