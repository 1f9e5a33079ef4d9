"""`pyLOM.DMD` entry points on the POD basis (pyLOM/DMD/__init__.py): run, frequency_damping, mode_computation,
reconstruction_jovanovic."""
from .wrapper import run, frequency_damping, mode_computation, reconstruction_jovanovic
