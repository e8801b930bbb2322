"""Import alias: ``meshrcnn_b200`` is the importable name of the package whose sources live in
``mesh_r-cnn_computer_vision_project_b200/`` (that directory name is not a legal Python identifier).
All sub-modules (``meshrcnn_b200.layers``, ``.loss_functions``, ``.mesh_sampling``, ...) resolve there.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                     "mesh_r-cnn_computer_vision_project_b200")
__path__ = [_SRC]
__version__ = "0.1.0"


def package_dir() -> str:
    return _SRC
