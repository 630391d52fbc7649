"""raytrace-miniapp_b200 — B200-native image-formation path of the XRayTrace CreateImage miniapp.

Import as `raytrace_miniapp_b200` (the shim package next to this directory maps the
importable name onto this hyphenated directory).
"""
from .abi import (BeamGrid, Gain, Problem, SeedProfile, N_SUB, N_MAX, K_MAX, ray_dtype,  # noqa: F401
                  OK, RAYS_FAILED, ERR_LIMITS, ERR_GRID, ERR_CUDA, ERR_ARG, ERR_FORMAT,
                  FLAG_NO_LIMITS, FLAG_LAZY_TABLES)
from .datfile import read_dat, write_dat, parse_payload, pack_payload  # noqa: F401

__version__ = "0.1.0"
