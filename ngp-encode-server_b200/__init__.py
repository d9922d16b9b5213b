"""ngp-encode-server_b200 -- B200 (sm_100a) implementation of ngp-encode-server's per-frame
pixel pipeline (unpack -> composite -> text overlay -> RGB/GRAY -> YUV420P) behind a C ABI.

The directory name carries a hyphen (it mirrors the reference's name); import it as
``ngp_encode_server_b200`` (alias package at the repo root).
"""
from .api import *  # noqa: F401,F403
from . import api  # noqa: F401
from . import synth  # noqa: F401,E402
from . import shard  # noqa: F401,E402
from . import avhandoff  # noqa: F401,E402
