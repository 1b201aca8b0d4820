"""Drop-in for libs/nerfheads/trainhead.py: `head.file B200NeRFHead`."""
import os
import sys

_REPO = os.environ.get("GPNERF_B200_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

import gpnerf_b200  # noqa: E402,F401
from gpnerf_b200.nerfhead import NeRFHead, NeRFRGBHead, NeRFSigmaHead, build_head  # noqa: E402,F401
