"""`pyLOM.POD` hot-path entry points (pyLOM/POD/__init__.py:9)."""
from .wrapper import run, truncate, reconstruct
from .utils import extract_modes
