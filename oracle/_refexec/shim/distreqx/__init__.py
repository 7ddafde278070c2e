"""Stand-in for distreqx 0.0.2 (uv.lock:385-386) -- TEST INFRASTRUCTURE, see oracle/_refexec/__init__.py.
Categorical / Normal / MultivariateNormalDiag as the reference's wrappers call them
(distribution/base_distribution.py:57-78, categorical.py:38, normal.py:38)."""

from . import bijectors, distributions  # noqa: F401
