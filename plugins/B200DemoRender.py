"""Drop-in for libs/renders/demo_render.py (progressive inference renderer).
`python tools/inference.py --cfg … render.file B200DemoRender` (README.md:73-79)."""
import os
import sys

_REPO = os.environ.get("GPNERF_B200_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

import gpnerf_b200  # noqa: E402,F401
from gpnerf_b200.render import Projector, Renderer  # noqa: E402,F401
from gpnerf_b200.render import build_render as _build  # noqa: E402


def build_render(cfg):
    return _build(cfg, progressive=True)
