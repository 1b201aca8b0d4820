"""Drop-in for libs/encoders/UNet.py: `encoder.file B200ResUNet` (train.py:143 / inference.py:61 look up
`build_encoder(cfg)` in that file).  Same `state_dict` as the reference's ResUNet, so checkpoints load unchanged."""
import os
import sys

_REPO = os.environ.get("GPNERF_B200_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _REPO not in sys.path:
    sys.path.insert(0, _REPO)

import gpnerf_b200  # noqa: E402,F401
from gpnerf_b200.encoder import ResUNet, build_encoder  # noqa: E402,F401
