"""An experiment build of the library: python tools/build_variant.py NAME -DFLAG[=V] ...  ->  build/NAME.so

Same sources and flags as momlevel_b200/_build.py plus the given ones; select it with MOMLEVEL_B200_LIB=build/NAME.so.
"""
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from momlevel_b200 import _build  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
out = ROOT / "build" / f"{name}.so"
out.parent.mkdir(exist_ok=True)
cmd = [_build._nvcc()] + _build.NVCC_FLAGS + extra + ["-o", str(out)] + [str(_build.CSRC / s) for s in _build.SOURCES]
res = subprocess.run(cmd, capture_output=True, text=True)
sys.stderr.write(res.stdout[-2000:] + res.stderr[-4000:])
print(out if res.returncode == 0 else "FAILED")
sys.exit(res.returncode)
