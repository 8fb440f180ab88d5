"""TEST INFRASTRUCTURE — re-export of the seeded synthetic weights (``plip_b200.synthetic``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline may import ``oracle``."""
from plip_b200.synthetic import (BOS, EOS, LOGIT_SCALE_INIT, PROJ, TEXT, VISION, make_state_dict)  # noqa: F401
