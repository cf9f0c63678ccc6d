"""pdlp-b200: B200-native hot path of OR-Tools PDLP behind the reference's surface.

The directory is named ``or-tools_b200`` (as the project layout requires); import
it as ``ortools_b200`` through the alias module at the repository root.
"""
from . import pdlp  # noqa: F401

__version__ = "0.1.0"
