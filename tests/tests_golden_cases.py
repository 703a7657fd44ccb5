"""Case table of tests/golden/make_golden_frame.py, importable from the tests."""
import importlib.util
import os

_spec = importlib.util.spec_from_file_location('make_golden_frame', os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden',
                                                                                  'make_golden_frame.py'))
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)
CASES = _mod.CASES
