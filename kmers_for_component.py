"""Top-level shim with the reference's module name: `from kmers_for_component import
kmers_for_component` (shannon.py:6,467) resolves to the B200 implementation."""
from shannon_b200.kmers_for_component import LAST_TIMINGS, kmers_for_component  # noqa: F401
