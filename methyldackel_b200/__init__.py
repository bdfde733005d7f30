"""B200-native `MethylDackel extract` / `mbias` hot path.

The product is two in-tree shared objects (built by ``__graft_entry__.build()``):

* ``lib/libmdgpu.so``  — hand-written CUDA (sm_100a) kernels behind the C ABI of ``include/mdgpu.h``;
* ``lib/libmdhost.so`` — host side (BAM decode to SoA tiles, reference chunk replay, text formatter,
  ``extract`` / ``mbias`` sub-command mains) behind ``include/mdhost.h``;

plus ``lib/MethylDackel``, the drop-in command-line binary.  This Python package is a thin
ctypes mirror used by the tests and ``bench.py``.
"""
from . import _abi  # noqa: F401
from .api import GpuContext, extract_main, mbias_main, read_region, fetch_contig  # noqa: F401
