"""diaglib_b200 — B200-native (sm_100a) replacement of the hot path of Molecolab-Pisa/diaglib:
the lobpcg_driver / davidson_driver iteration body and the block kernels under it.

The product is the C-ABI shared library ``libdiaglib_b200.so`` (include/diaglib_b200.h).  This
package is the thin host-side mirror of the reference's interface (same routine names,
argument meaning and error behaviour) over ctypes.  There is no CPU fallback: if the library
or a CUDA device is missing, calls fail loudly.
"""
from .api import (  # noqa: F401
    DiaglibError,
    build,
    davidson_driver,
    gen_david_driver,
    caslr_eff_driver,
    set_lr,
    init,
    lib,
    lib_path,
    lobpcg_driver,
    b_ortho,
    b_ortho_vs_x,
    ortho,
    ortho_cd,
    ortho_vs_x,
    set_csr,
    set_csr_b,
    set_csr_device,
    set_csr_row_order,
    set_halo,
    last_history,
    last_stats,
    peer_info,
    last_timers,
    set_profile,
)
