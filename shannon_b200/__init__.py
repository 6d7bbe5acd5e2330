"""shannon_b200: B200-native k-mer front end of the Shannon RNA-Seq assembler.

Host-side mirror of the reference's two hot-path modules (same entry points, arguments and output
files) on top of a C-ABI library of hand-written sm_100a CUDA kernels (include/shannon_b200.h).
No CPU fallback: importing works anywhere, computing needs the built library and a GPU.
"""
__all__ = ["extension_correction", "kmers_for_component", "get_context"]


def __getattr__(name):
    if name == "get_context":
        from .extension_correction import get_context
        return get_context
    raise AttributeError(name)
