"""Re-export of the seeded synthetic weights/inputs (they live in the package so that bench.py's
product arm does not import oracle/)."""
from e2e_multi_view_matching_b200.synthetic import *  # noqa: F401,F403
from e2e_multi_view_matching_b200.synthetic import BN_EPS  # noqa: F401
