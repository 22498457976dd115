"""Launcher used by tests/test_reference_cli.py: executes an UNMODIFIED reference script (path in argv[1]) with the repo's drop-in
modules first on sys.path (so `from pix2pix_turbo import Pix2Pix_Turbo` / `from cyclegan_turbo import CycleGAN_Turbo` resolve to
img2img-turbo_b200/) and the reference's own src/ directory after it (for image_prep / my_utils, which stay the reference's).
On a GPU-less host the CPU test double of the engine is installed first (tests/cpu_stub_engine.py)."""
import json
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
script = sys.argv[1]
sys.argv = [script] + sys.argv[2:]
sys.path[:0] = [os.path.join(ROOT, "img2img-turbo_b200"), os.path.dirname(os.path.abspath(script))]

import torch  # noqa: E402

if not torch.cuda.is_available():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cpu_stub_engine
    cpu_stub_engine.install()
runpy.run_path(script, run_name="__main__")
if not torch.cuda.is_available():
    print("STUB_CALLS " + json.dumps([list(map(str, c)) for c in cpu_stub_engine.CALLS]))
import pix2pix_turbo  # noqa: E402  (proves whose module the script imported)
print("MODULE_FILE " + os.path.abspath(pix2pix_turbo.__file__))
