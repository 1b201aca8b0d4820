"""gpnerf_b200 – B200-native (sm_100a) implementation of GP-NeRF's
geometry-guided progressive volume-rendering hot path behind the reference's
``build_render`` / ``build_head`` plugin API.  See DESIGN.md."""
__version__ = "0.1.0"
