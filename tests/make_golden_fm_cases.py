"""The case list of tests/golden/make_golden_fm.py, importable without cv2 (the generator needs it,
the consumers of the committed .npz do not)."""
import os

_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden_fm.py")
_src = open(_path).read()
_start = _src.index("CASES = [")
_end = _src.index("]\n", _start) + 1
_ns = {}
exec(_src[_start:_end], _ns)
CASES = _ns["CASES"]
