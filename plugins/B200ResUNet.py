"""Drop into the reference's libs/encoders/ and set `encoder.file: B200ResUNet` in the experiment yaml
(train.py:143 / inference.py:61 look up `build_encoder(cfg)` in that file).  Same `state_dict` as
libs/encoders/UNet.py, so existing checkpoints load unchanged."""
from gpnerf_b200.encoder import ResUNet, build_encoder  # noqa: F401
