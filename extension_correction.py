"""Top-level shim with the reference's module name: `from extension_correction import
extension_correction` (shannon.py:5,459) resolves to the B200 implementation."""
from shannon_b200.extension_correction import *  # noqa: F401,F403
from shannon_b200.extension_correction import LAST_TIMINGS, extension_correction, run_correction  # noqa: F401

if __name__ == '__main__':
    import sys
    extension_correction(sys.argv[1:] if len(sys.argv) > 1 else
                         ['kmers.dict', 'allowed_kmers.dict', '1', '1', '-d'])
