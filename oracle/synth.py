"""TEST INFRASTRUCTURE — re-export of the seeded synthetic inputs (``plip_b200.synthetic``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU baseline may import ``oracle``."""
from plip_b200.synthetic import pixel_values, tiles_u8, token_ids  # noqa: F401
