"""Puts the repository root on ``sys.path`` so the bare-name shim modules next to this file can import the package."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
