"""mumemto_b200 -- B200-native multi-MUM / multi-MEM engine (CUDA sm_100a behind a C ABI, include/mumemto_b200.h).

Host-side Python mirror of the reference's library interface (python_bindings/src/mumemto_pybind.cpp:93-118):
``mum(sequences, ...)`` and ``mem(sequences, ...)``.  Everything runs on the GPU; importing works anywhere, but using the
engine without the built CUDA library or without a device raises.
"""
from ._binding import Engine, MumemtoError, LIB_PATH, read_fasta_image  # noqa: F401
from .api import mum, mem, MumResult, MemResult, find_files, write_outputs, merge_files, read_mums  # noqa: F401

__version__ = "0.1.0"
